import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eeyore_b200 import stats as st
gd = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/stats_goldens.npz"))
x = torch.from_numpy(gd["chains"])
out = st.chain_stats(x, want=("mean", "inse"))
for i in range(4):
    m = out["inse"][i]
    print(i, "sym", torch.equal(m, m.t()), (m - m.t()).abs().max().item())
w = out["inse"].mean(0)
print("w sym", torch.equal(w, w.t()), (w - w.t()).abs().max().item(), torch.linalg.cholesky_ex(w).info)
print(w)
