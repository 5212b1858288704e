"""GPU check of the tcgen05 data-parallel kernel against the FP32 CUDA-core kernel and the numpy oracle, with timings.
   python tools/check_dp_tc.py [n_rows_for_timing]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402
from eeyore_b200 import _native as nv  # noqa: E402
from oracle.mlp import MLPSpec, BINARY  # noqa: E402

P = 5313
SPEC = MLPSpec([16, 64, 64, 1], loss=BINARY)


def sums(fn, theta, x, y, ws=None):
    out = torch.empty(P + 1, dtype=torch.float64, device="cuda")
    nv.check(fn(nv.ptr(theta), nv.ptr(x), nv.ptr(y), x.shape[0], nv.ptr(out), nv.ptr(ws) if ws is not None else None, None))
    return out


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def main():
    lib = nv.lib()
    tc, ff = lib.eeyore_b200_dp_loglik_grad, lib.eeyore_b200_dp_loglik_grad_ffma
    ok = True
    for n in (1, 5, 127, 128, 129, 1000, 4099, 40000):
        rng = np.random.default_rng(n)
        x = rng.normal(size=(n, 16)).astype(np.float32)
        t = rng.normal(size=16).astype(np.float32)
        y = ((x @ t + 0.5 * rng.normal(size=n)) > 0).astype(np.float32)
        theta = (rng.normal(size=P) * 0.3).astype(np.float32)
        xd, yd, td = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(theta).cuda()
        a = sums(tc, td, xd, yd).cpu().numpy()
        b = sums(ff, td, xd, yd).cpu().numpy()
        th64 = theta.astype(np.float64)[None]
        ll_ref = oracle.log_lik(SPEC, th64, x.astype(np.float64), y[:, None])
        _, g_ref = oracle.log_target_grad(SPEC, th64, x.astype(np.float64), y[:, None], np.zeros(P), np.full(P, 3 ** 0.5))
        g_ref = g_ref + th64 / 3.0          # remove the Normal(0, sqrt 3) prior term: gradient of the log-likelihood
        line = f"n={n:6d}  tc vs ffma: ll {abs(a[0]-b[0])/abs(b[0]):.2e} grad {rel(a[1:], b[1:]):.2e}"
        if ll_ref is not None:
            line += f" | tc vs oracle: ll {abs(a[0]-ll_ref[0])/abs(ll_ref[0]):.2e} grad {rel(a[1:], g_ref[0]):.2e}" \
                    f" | ffma vs oracle: grad {rel(b[1:], g_ref[0]):.2e}"
            ok &= rel(a[1:], g_ref[0]) < 1e-5
        else:
            ok &= rel(a[1:], b[1:]) < 5e-6
        # per-block errors (W0, b0, W1, b1, W2, b2)
        blocks = {"W0": (1, 1025), "b0": (1025, 1089), "W1": (1089, 5185), "b1": (5185, 5249), "W2": (5249, 5313), "b2": (5313, 5314)}
        line += "  [" + " ".join(f"{k}:{rel(a[lo:hi], b[lo:hi]):.1e}" for k, (lo, hi) in blocks.items()) + "]"
        print(line, flush=True)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(n, 16, device="cuda", generator=g)
    y = (torch.rand(n, device="cuda", generator=g) < 0.5).float()
    theta = torch.randn(P, device="cuda", generator=g) * 0.1
    ws = torch.empty(lib.eeyore_b200_dp_workspace_bytes() // 8, dtype=torch.float64, device="cuda")
    amax = torch.zeros(1, dtype=torch.float32, device="cuda")
    nv.check(lib.eeyore_b200_dp_absmax(nv.ptr(x), x.numel(), nv.ptr(amax), None))

    def tc_scaled(th, xx, yy, nn, out, wsp, st):      # max |x| computed once, as DataShardedHMC does
        return lib.eeyore_b200_dp_loglik_grad_x(th, xx, yy, nn, nv.ptr(amax), out, wsp, st)

    for name, fn in (("tcgen05", tc_scaled), ("ffma", ff)):
        for _ in range(3):
            sums(fn, theta, x, y, ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            sums(fn, theta, x, y, ws)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:8s}: {ms:.3f} ms per evaluation of {n} rows = {n * 29056 / ms / 1e9:.1f} TFLOP/s algorithmic", flush=True)
    print("CHECK", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
