"""Accuracy bounds of the fp64 fast math (table exp / sigmoid / log, branch-free sqrt and sincos, Box-Muller edge cases) that
the chain kernels run, checked on the host build of the very same __host__ __device__ code (tests/hostsim/fastmath_check.cpp
against long-double libm).  The parity bar of the path is 1e-10 relative; these bounds are five to six orders tighter."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def report():
    src = ROOT / "tests" / "hostsim" / "fastmath_check.cpp"
    exe = ROOT / "tests" / "hostsim" / "fastmath_check.bin"
    deps = [src] + list((ROOT / "eeyore_b200" / "csrc").glob("*.cuh")) + list((ROOT / "eeyore_b200" / "csrc").glob("*.inc"))
    if not exe.exists() or exe.stat().st_mtime < max(d.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-I" + str(ROOT / "eeyore_b200" / "csrc"), "-I/usr/local/cuda/include",
                        "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    return {k: float(v) for k, v in (line.split() for line in out.strip().splitlines())}


def test_sigmoid_exp_log_accuracy(report):
    assert report["sigmoid_rel_small"] < 8e-16            # |g| <= 3
    assert report["sigmoid_rel_scaled"] < 1.0             # 6e-16 + 1.2e-16 |g| everywhere (|g| 2^-53 from the base-2 reduction)
    assert report["exp_rel_scaled"] < 1.0
    assert report["log_rel"] < 4e-16 and report["log_abs_near_one"] < 5e-18   # |log x| <= 1e-3: absolute bound
    assert report["log_one"] == 0.0 and report["log_min_normal_err"] < 3e-13


def test_sigmoid_saturation_and_nan(report):
    assert report["sig_40_is_one"] == 1 and report["sig_inf"] == 1.0 and report["sig_big"] == 1.0
    assert report["sig_minf"] == 0.0 and report["sig_mbig"] == 0.0 and report["sig_nan_is_nan"] == 1


def test_box_muller_pieces(report):
    assert report["sin_abs"] < 3e-16 and report["cos_abs"] < 3e-16 and report["sqrt_rel"] < 3e-16
    assert report["umax_is_one"] == 1 and report["bm_umax_abs"] == 0.0     # radius exactly 0, not NaN
    assert report["bm_umin_err"] < 1e-14
