"""ctypes binding of libeeyore_b200.so (C ABI in include/eeyore_b200.h) and the in-tree build recipe.

There is no CPU fallback: every compute entry point raises RuntimeError when the shared library is missing or
no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import torch

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_PATH = PKG / "libeeyore_b200.so"
BUILD_DIR = PKG / "_build"

F32, F64 = 0, 1
ACT_NONE, ACT_SIGMOID = 0, 1
LOSS_BINARY, LOSS_MULTICLASS = 0, 1
RNG_PHILOX, RNG_TAPE = 0, 1
EINVAL, EUNSUPPORTED, ECUDA, ENUMERIC = -1, -2, -3, -4

DTYPE_IDS = {torch.float32: F32, torch.float64: F64}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


class RunParams(C.Structure):
    """struct eeyore_b200_run_params"""
    _fields_ = [
        ("n_chains", C.c_int64), ("n_iters", C.c_int64), ("n_burnin", C.c_int64), ("thin", C.c_int64),
        ("step", C.c_double), ("num_steps", C.c_int32), ("symmetric", C.c_int32),
        ("has_temperature", C.c_int32), ("rng_mode", C.c_int32), ("temperature", C.c_double),
        ("seed", C.c_uint64), ("iter_offset", C.c_uint64), ("chain_offset", C.c_uint64),
        ("z_tape", C.c_void_p), ("u_tape", C.c_void_p),
        ("x", C.c_void_p), ("y", C.c_void_p), ("n_rows", C.c_int64),
        ("prior_loc", C.c_void_p), ("prior_scale", C.c_void_p),
        ("theta", C.c_void_p), ("target", C.c_void_p), ("grad", C.c_void_p),
        ("st_chain", C.c_int64), ("st_param", C.c_int64),
        ("out_samples", C.c_void_p), ("ss_iter", C.c_int64), ("ss_chain", C.c_int64), ("ss_param", C.c_int64),
        ("out_target", C.c_void_p), ("out_grad", C.c_void_p), ("out_accepted", C.c_void_p),
        ("accept_count", C.c_void_p),
        ("lanes_per_chain", C.c_int32), ("reserved", C.c_int32),
        ("tuner_l", C.c_double), ("tuner_d", C.c_double), ("tuner_m", C.c_double), ("tuner_logeub", C.c_double),
        ("tuner_has_eub", C.c_int32), ("tuner_pad", C.c_int32), ("tuner_iter0", C.c_int64), ("tuner_burnin", C.c_int64),
        ("tuner_state", C.c_void_p),
        ("stream", C.c_void_p),
        ("adapt_p", C.c_double * 3), ("adapt_t0", C.c_int32), ("adapt_pad", C.c_int32), ("adapt_iter0", C.c_int64),
        ("adapt_state", C.c_void_p), ("adapt_cov0", C.c_void_p), ("adapt_status", C.c_void_p),
        ("final_theta", C.c_void_p), ("fs_chain", C.c_int64), ("fs_param", C.c_int64),
        ("final_target", C.c_void_p), ("final_accept_count", C.c_void_p),
    ]


# symbol -> (restype, argtypes); must list every function declared in include/eeyore_b200.h
_VP, _I, _I64, _U64, _D = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
SIGNATURES = {
    "eeyore_b200_last_error": (C.c_char_p, []),
    "eeyore_b200_version": (C.c_char_p, []),
    "eeyore_b200_mlp_create": (_I, [_I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), _I, _I, C.POINTER(_VP)]),
    "eeyore_b200_mlp_destroy": (_I, [_VP]),
    "eeyore_b200_mlp_num_params": (_I, [_VP]),
    "eeyore_b200_mlp_is_specialised": (_I, [_VP]),
    "eeyore_b200_log_target_grad": (_I, [_VP, _I64, _VP, _VP, _VP, _I64, _VP, _VP, _I, _D, _VP, _VP, _VP, _VP, _I, _VP]),
    "eeyore_b200_forward": (_I, [_VP, _I64, _VP, _VP, _I64, _VP, _VP]),
    "eeyore_b200_num_saved": (_I64, [_I64, _I64, _I64]),
    "eeyore_b200_mh_run": (_I, [_VP, C.POINTER(RunParams)]),
    "eeyore_b200_mala_run": (_I, [_VP, C.POINTER(RunParams)]),
    "eeyore_b200_hmc_run": (_I, [_VP, C.POINTER(RunParams)]),
    "eeyore_b200_smmala_run": (_I, [_VP, C.POINTER(RunParams)]),
    "eeyore_b200_am_run": (_I, [_VP, C.POINTER(RunParams)]),
    "eeyore_b200_ram_run": (_I, [_VP, C.POINTER(RunParams)]),
    "eeyore_b200_adapt_state_len": (_I64, [_VP, _I]),
    "eeyore_b200_chain_stats": (_I, [_I, _I64, _I64, _I, _VP, _I64, _I64, _I64, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP, _VP, _I, _VP]),
    "eeyore_b200_dp_num_params": (_I, []),
    "eeyore_b200_dp_loglik_grad": (_I, [_VP, _VP, _VP, _I64, _VP, _VP, _VP]),
    "eeyore_b200_dp_loglik_grad_x": (_I, [_VP, _VP, _VP, _I64, _VP, _VP, _VP, _VP]),
    "eeyore_b200_dp_absmax": (_I, [_VP, _I64, _VP, _VP]),
    "eeyore_b200_dp_loglik_grad_ffma": (_I, [_VP, _VP, _VP, _I64, _VP, _VP, _VP]),
    "eeyore_b200_dp_num_parts": (_I, [_I64]),
    "eeyore_b200_dp_post": (_I, [_VP, _I, _I, _I, _U64, C.POINTER(_VP), _VP, _VP, _VP, _VP, _I, _D, _VP, _VP, _I, _D, _VP, _VP,
                                 _VP, _VP, _VP]),
    "eeyore_b200_dp_exchange_bytes": (_I64, []),
    "eeyore_b200_dp_exchange_create": (_I, [C.POINTER(_VP), _VP]),
    "eeyore_b200_dp_exchange_open": (_I, [_VP, C.POINTER(_VP)]),
    "eeyore_b200_dp_exchange_close": (_I, [_VP]),
    "eeyore_b200_dp_exchange_destroy": (_I, [_VP]),
    "eeyore_b200_dp_exchange_scratch_offset": (_I64, []),
    "eeyore_b200_dp_scratch_len": (_I64, []),
    "eeyore_b200_dp_hmc_run": (_I, [_VP, _VP, _I64, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _D, _D, _I,
                                    _I64, _I64, _U64, _U64, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _U64, C.POINTER(_VP), _VP]),
    "eeyore_b200_dp_workspace_bytes": (_I64, []),
    "eeyore_b200_dp_finish": (_I, [_VP, _VP, _VP, _VP, _I, _D, _VP, _VP, _VP]),
    "eeyore_b200_dp_hmc_begin": (_I, [_VP, _VP, _D, _U64, _U64, _VP, _VP, _VP, _VP, _VP]),
    "eeyore_b200_dp_hmc_step": (_I, [_VP, _D, _I, _VP, _VP, _VP, _VP]),
    "eeyore_b200_dp_hmc_accept": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _U64, _U64, _VP, _VP, _VP, _VP, _VP, _VP]),
    "eeyore_b200_philox_draws": (_I, [_I, _I64, _I, _U64, _U64, _U64, _VP, _VP, _VP]),
    "eeyore_b200_fma_peak": (_I, [_I, _I, C.POINTER(_D)]),
}

_lib = None


def sources():
    return sorted(CSRC.glob("*.cu"))


def build(verbose=False, jobs=None, extra_flags=(), out=None, build_dir=None):
    """Compile every CUDA source for sm_100a and link libeeyore_b200.so in-tree (nvcc cross-compiles without a GPU).
    extra_flags / out / build_dir build an experimental variant next to the default library (EEYORE_B200_LIB selects it)."""
    global BUILD_DIR, LIB_PATH
    saved = (BUILD_DIR, LIB_PATH)
    if out is not None:
        LIB_PATH = Path(out)
        BUILD_DIR = Path(build_dir or (str(out) + ".build"))
    try:
        return _build(verbose, jobs, tuple(extra_flags))
    finally:
        BUILD_DIR, LIB_PATH = saved


def _build(verbose, jobs, extra_flags):
    BUILD_DIR.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "eeyore_b200.h"]
    newest_header = max(h.stat().st_mtime for h in headers)
    objs, todo = [], []
    for src in sources():
        obj = BUILD_DIR / (src.stem + ".o")
        objs.append(obj)
        if not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, newest_header):
            todo.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = ["nvcc", *NVCC_FLAGS, *extra_flags, "-c", str(src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    if todo:
        with ThreadPoolExecutor(max_workers=jobs or min(len(todo), os.cpu_count() or 4)) as ex:
            list(ex.map(compile_one, todo))
    if todo or not LIB_PATH.exists():
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


def lib():
    """The loaded shared library (raises RuntimeError if it has not been built)."""
    global _lib
    if _lib is None:
        path = Path(os.environ.get("EEYORE_B200_LIB", LIB_PATH))
        if not path.exists():
            raise RuntimeError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(eeyore_b200 has no CPU fallback)")
        l = C.CDLL(str(path))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("eeyore_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def check(rc):
    if rc == 0:
        return
    msg = lib().eeyore_b200_last_error().decode()
    if rc in (EINVAL, EUNSUPPORTED):
        raise ValueError(msg)
    raise RuntimeError(msg)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
