"""Pinned host <-> device copy bandwidth of this box (what bounds bench.py's e2e: every saved state goes back to the host)."""
import json
import torch

n = 512 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    out[name + "_gbs"] = 5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
# both directions at once on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
torch.cuda.current_stream().wait_stream(s1)
torch.cuda.current_stream().wait_stream(s2)
e1.record()
torch.cuda.synchronize()
out["duplex_each_gbs"] = 5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
print(json.dumps(out))
