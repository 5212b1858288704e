"""Shared helpers for tests: golden loading and architecture table."""
from pathlib import Path

import numpy as np

from oracle.mlp import MLPSpec, BINARY, MULTICLASS

GOLDEN = Path(__file__).resolve().parent / "golden"

ARCHS = {
    "221": dict(dims=[2, 2, 1], loss=BINARY, data="xor"),
    "2321": dict(dims=[2, 3, 2, 1], loss=BINARY, data="xor"),
    "433": dict(dims=[4, 3, 3], loss=MULTICLASS, data="iris"),
    "4323": dict(dims=[4, 3, 2, 3], loss=MULTICLASS, data="iris"),
}
PRIOR_SCALES = {"p1": 1.0, "p100": 100.0, "psqrt3": 3.0 ** 0.5}
NP_DTYPES = {"f64": np.float64, "f32": np.float32}
# tolerances stated by BASELINE.json north_star: 1e-10 relative at fp64, 1e-5 at fp32
RTOL = {"f64": 1e-10, "f32": 1e-5}


def spec_of(arch):
    a = ARCHS[arch]
    return MLPSpec(dims=a["dims"], loss=a["loss"])


def load(name):
    return np.load(GOLDEN / f"{name}.npz")


def data_of(arch, dtype, mg=None):
    mg = mg if mg is not None else load("model_goldens")
    d = ARCHS[arch]["data"]
    return mg[f"{d}_x"].astype(dtype), mg[f"{d}_y"].astype(dtype)


def rel_err(a, b):
    """max |a-b| / max(|b|) -- vector-relative error (scale of the reference vector)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)


def pp_setup(name):
    """Power-posterior golden -> (spec, x, y, kinds, kwargs, arrays) in the oracle's calling convention."""
    gd = load(name)
    arch = "221" if "221" in name else "2321"
    spec = spec_of(arch)
    x, y = data_of(arch, np.float64)
    kinds = ["mh" if str(k) == "MetropolisHastings" else "mala" for k in gd["kinds"]]
    kwargs = [({} if k == "mh" else {"step": float(s)}) for k, s in zip(kinds, gd["steps"])]
    return gd, spec, x, y, kinds, kwargs
