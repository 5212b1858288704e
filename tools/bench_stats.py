"""Times the diagnostics kernel alone on synthetic chains (one GPU).
    python tools/bench_stats.py [--chains 65536] [--n 1000] [--p 20] [--rho 0.0] [--layout cnp|npc] [--max-lag 10] [--reps 5]
rho = AR(1) coefficient of the synthetic chains (0 = white noise: the INSE loop stops after one or two lag pairs)."""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eeyore_b200 import stats as st  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=65536)
ap.add_argument("--n", type=int, default=1000)
ap.add_argument("--p", type=int, default=20)
ap.add_argument("--rho", type=float, default=0.0)
ap.add_argument("--layout", default="cnp")
ap.add_argument("--max-lag", type=int, default=10)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--want", default="ess")
a = ap.parse_args()
dt = torch.float64 if a.dtype == "f64" else torch.float32
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(a.chains, a.n, a.p, dtype=dt, device="cuda", generator=g)
if a.rho:
    for t in range(1, a.n):
        x[:, t] += a.rho * x[:, t - 1]
if a.layout == "npc":
    x = x.permute(1, 2, 0).contiguous()
want = tuple(w for w in a.want.split(",") if w)
ml = a.max_lag if a.max_lag >= 0 else None
out = st.chain_stats(x, layout=a.layout, want=want, max_lag=ml, check=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    out = st.chain_stats(x, layout=a.layout, want=want, max_lag=ml, check=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
lags = out["lags"].double()
passes = 1 + (1 if ml is not None else 0) + (1 + (lags[:, 1].clamp(min=0) + 1).mean().item() if want else 0)
flops = 2.0 * a.n * a.p * a.p * (2 + (lags[:, 1].clamp(min=0) + 1).mean().item()) * a.chains if want else 0.0
print(json.dumps({"chains": a.chains, "n": a.n, "P": a.p, "rho": a.rho, "layout": a.layout, "ms": ms,
                  "chains_per_s": a.chains / ms * 1e3, "mean_m_last": lags[:, 1].mean().item(), "max_m_last": lags[:, 1].max().item(),
                  "status_bad": int((out["status"] != 0).sum()), "approx_passes": passes,
                  "lag_product_tflops": flops / ms / 1e9,
                  "hbm_gbs_if_every_pass_streams": passes * a.chains * a.n * a.p * x.element_size() / ms / 1e6}))
