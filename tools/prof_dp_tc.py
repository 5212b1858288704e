"""Phase-level cycle profile of the tcgen05 data-parallel kernel (needs a -DDP_TC_PROFILE build):
   python tools/prof_dp_tc.py   (builds eeyore_b200/libeeyore_b200_prof.so if missing; run on the GPU box)"""
import ctypes as C
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
LIB = ROOT / "eeyore_b200" / "libeeyore_b200_prof.so"
if "build" in sys.argv:
    from eeyore_b200 import _native as nv
    nv.build(extra_flags=["-DDP_TC_PROFILE"], out=LIB)
    sys.exit(0)
os.environ["EEYORE_B200_LIB"] = str(LIB)
import torch  # noqa: E402
from eeyore_b200 import _native as nv  # noqa: E402

P = 5313
NAMES = ["loop top", "P0 split+sync", "wait MMA1", "P1 + sync", "wait MMA2", "P2 + sync", "wait MMA3", "P3 compute", "wait MMA4",
         "P3 store + sync", "P4 dW1 fold", "wait MMA5", "P4 dW0 fold"]
lib = nv.lib()
n = 1 << 21
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, 16, device="cuda", generator=g)
y = (torch.rand(n, device="cuda", generator=g) < 0.5).float()
theta = torch.randn(P, device="cuda", generator=g) * 0.1
out = torch.empty(P + 1, dtype=torch.float64, device="cuda")
ws = torch.empty(lib.eeyore_b200_dp_workspace_bytes() // 8, dtype=torch.float64, device="cuda")
buf = (C.c_ulonglong * 24)()
prof = lib.eeyore_b200_dp_tc_profile
prof.argtypes = [C.POINTER(C.c_ulonglong)]
for it in range(3):
    nv.check(lib.eeyore_b200_dp_loglik_grad(nv.ptr(theta), nv.ptr(x), nv.ptr(y), n, nv.ptr(out), nv.ptr(ws), None))
    prof(buf)
tiles = (n // 128 + 147) // 148
labels = {1: "P0 (both contexts) + MMA1 issue", 3: "P1 (wait MMA1, sigmoid, split) x2", 5: "P2 (wait MMA2, head, Delta2, logs, butterfly) x2",
          9: "P3 (wait MMA3/4, Delta1) x2", 12: "fold (every 2 pairs)"}
tot = sum(buf[i] for i in labels)
print(f"CTA 0: {tiles} tiles, {tot / tiles:.0f} cycles per tile; prologue {buf[20]} cycles, after the loop {buf[21]} cycles")
for i, nm in labels.items():
    print(f"  {nm:52s} {buf[i] / tiles:8.0f} per tile  {100 * buf[i] / tot:5.1f}%")
