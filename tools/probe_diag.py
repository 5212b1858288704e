"""Distribution of the INSE loop lengths (sn, m_last) over real HMC chains of BASELINE config 4, and the stage times."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from eeyore_b200 import stats as st
from eeyore_b200.chains import ChainList

class A: pass
args = A(); args.gpus = 1; args.chains = 0; args.iters = 0; args.lanes = 0
ctx = bench.Ctx(args)
w = bench.WORKLOADS["cfg4"]
model, loader, dt, x, y = bench.build_model(ctx, w)
C, n = 65536, 1000
theta = torch.randn(C, model.num_params(), dtype=dt, generator=torch.Generator().manual_seed(2000))
s = bench.make_sampler(w, model, loader, theta, 4242, 1, 0, chain=ChainList(keys=["sample", "accepted"]))
s.sample_layout = "cnp"
s.run(num_epochs=n, num_burnin_epochs=0)
ring = s._device_blocks[-1]["sample"]
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = st.chain_stats(ring, layout="npc", want=("ess",), max_lag=10, check=False)
    e1.record(); torch.cuda.synchronize()
    print("stats ms", e0.elapsed_time(e1))
lags = out["lags"].cpu().numpy(); status = out["status"].cpu().numpy()
acc = s._device_blocks[-1]["accepted"].float().mean(0).cpu().numpy()
print("status bad", (status != 0).sum(), "of", C)
print("sn hist", np.bincount(np.minimum(lags[:, 0], 20))[:21])
print("m_last hist", np.bincount(np.minimum(np.maximum(lags[:, 1], 0), 40))[:41])
bad = status != 0
print("acceptance of bad chains: mean %.3f min %.3f max %.3f; of good %.3f" % (acc[bad].mean(), acc[bad].min(), acc[bad].max(), acc[~bad].mean()))
sam = ring.permute(2, 0, 1)[torch.from_numpy(np.nonzero(bad)[0][:3]).cuda()].cpu().numpy()
for c in sam:
    print("bad chain column std:", np.round(c.std(0), 4))
