"""Bundled data sets (mirror of eeyore/datasets/data_info.py): XOR and Fisher's iris."""
from pathlib import Path

data_paths = {name: Path(__file__).resolve().parent.parent / "data" / name for name in ("iris", "xor")}
