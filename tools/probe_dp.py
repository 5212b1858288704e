import os, time, torch, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
from torch.distributions import Normal
from eeyore_b200.constants import loss_functions
from eeyore_b200.models.mlp import MLP, Hyperparameters
from eeyore_b200.samplers import DataShardedHMC, shard_rows
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
n_total = 8388608
lo, hi = shard_rows(n_total, world, rank)
g = torch.Generator(device=dev).manual_seed(4)
teacher = torch.randn(16, device=dev, generator=g)
x = torch.randn(n_total, 16, device=dev, generator=g); y = ((x @ teacher + 0.5*torch.randn(n_total, device=dev, generator=g)) > 0).float()
x, y = x[lo:hi].contiguous(), y[lo:hi].contiguous()
hp = Hyperparameters([16,64,64,1], 3*[True], 3*[torch.sigmoid])
m = MLP(loss=loss_functions["binary_classification"], hparams=hp, dtype=torch.float32, device=dev)
P = m.num_params(); m.prior = Normal(torch.zeros(P), 3**0.5*torch.ones(P))
th = (torch.randn(P, generator=torch.Generator().manual_seed(5)) * 0.1).pin_memory()
if world == 1:
    for step in (1e-4, 5e-5, 2e-5, 1e-5):
        s = DataShardedHMC(m, th, x, y, step=step, num_steps=10, seed=7)
        s.run(num_epochs=12, num_burnin_epochs=0); torch.cuda.synchronize()
        print("step", step, "accepted", s.acceptance_count(), "of 12", "lt", s._lt_c.item(), flush=True)
def timed(label, fn, n=3):
    torch.cuda.synchronize(); 
    if world > 1: dist.barrier()
    t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / n
    if rank == 0: print(f"{label}: {dt*1e3:.1f} ms", flush=True)
s = DataShardedHMC(m, th, x, y, step=2e-5, num_steps=10, seed=7)
timed("resident run(2)", lambda: s.run(num_epochs=2, num_burnin_epochs=0))
def fresh():
    s2 = DataShardedHMC(m, th, x, y, step=2e-5, num_steps=10, seed=7)
    s2.run(num_epochs=2, num_burnin_epochs=0)
timed("fresh sampler + run(2)", fresh)
def fresh_sync():
    s2 = DataShardedHMC(m, th, x, y, step=2e-5, num_steps=10, seed=7)
    r = s2.run(num_epochs=2, num_burnin_epochs=0); torch.cuda.current_stream().synchronize()
timed("fresh sampler + run(2) + sync", fresh_sync)
timed("ctor only", lambda: DataShardedHMC(m, th, x, y, step=2e-5, num_steps=10, seed=7))
import eeyore_b200._native as nv
sums = torch.empty(P+1, dtype=torch.float64, device=dev); thd = th.to(dev)
timed("dp_loglik_grad only x20", lambda: [nv.check(nv.lib().eeyore_b200_dp_loglik_grad(nv.ptr(thd), nv.ptr(x), nv.ptr(y), x.shape[0], nv.ptr(sums), None, nv.stream_ptr(dev))) for _ in range(20)])
if world > 1:
    timed("allreduce only x20", lambda: [dist.all_reduce(sums) for _ in range(20)])
    dist.destroy_process_group()
