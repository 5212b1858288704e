"""Multi-GPU check of the data-sharded HMC exchange (run under torchrun, one rank per GPU):
   * the peer-store exchange (exchange="p2p") and the NCCL all-reduce formulation give the same chain
   * every rank holds the same chain
   * the chain equals the one-shard run of rank 0 over the whole data set up to fp32 summation order
   * the persistent run kernel (exchange inside one cooperative launch) equals the per-evaluation launches bitwise
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_dp_p2p.py
Also run by tests/test_gpu_datapar.py::test_two_rank_peer_store_exchange when two GPUs are visible."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
from torch.distributions import Normal

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from eeyore_b200.constants import loss_functions  # noqa: E402
from eeyore_b200.models.mlp import MLP, Hyperparameters  # noqa: E402
from eeyore_b200.samplers import DataShardedHMC, shard_rows  # noqa: E402

P = 5313


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 300_007
    rng = np.random.default_rng(0)
    x = rng.normal(size=(n, 16)).astype(np.float32)
    t = rng.normal(size=16).astype(np.float32)
    y = ((x @ t + 0.5 * rng.normal(size=n)) > 0).astype(np.float32)
    theta0 = torch.from_numpy((rng.normal(size=P) * 0.1).astype(np.float32))
    hp = Hyperparameters([16, 64, 64, 1], 3 * [True], 3 * [torch.sigmoid])
    model = MLP(loss=loss_functions["binary_classification"], hparams=hp, dtype=torch.float32, device=dev)
    model.prior = Normal(torch.zeros(P), 3 ** 0.5 * torch.ones(P))
    lo, hi = shard_rows(n, world, rank)
    xs, ys = torch.from_numpy(x[lo:hi]).to(dev), torch.from_numpy(y[lo:hi]).to(dev)
    T, L, step = 6, 5, 2e-4
    out = {}
    for mode in ("p2p", "p2p-launches", "nccl"):
        s = DataShardedHMC(model, theta0, xs, ys, step=step, num_steps=L, seed=3, exchange=mode.split("-")[0],
                           trajectory="launches" if mode.endswith("launches") else "persistent")
        assert s.persistent == (mode == "p2p")
        samples, targets, accepted = s.run(num_epochs=T, num_burnin_epochs=0)
        torch.cuda.synchronize()
        s.check_status()
        out[mode] = (samples.clone(), targets.clone(), accepted.clone())
        # timing: evaluations per second of the trajectory loop
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.run(num_epochs=10, num_burnin_epochs=10)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            print(f"{mode}: {10 * L / dt:.1f} evaluations/s over {n} rows on {world} GPUs ({1e3 * dt / (10 * L):.3f} ms per evaluation)",
                  flush=True)
        s.close()
    ok = True
    # the persistent kernel (one cooperative launch per run, exchange inside) and the per-evaluation launches: bitwise equal
    same_traj = all(torch.equal(a, b) for a, b in zip(out["p2p"], out["p2p-launches"]))
    same = torch.equal(out["p2p"][0], out["nccl"][0]) and torch.equal(out["p2p"][2], out["nccl"][2])
    err_modes = (out["p2p"][0] - out["nccl"][0]).abs().max().item()
    # every rank holds the same chain
    ref = out["p2p"][0].clone()
    dist.broadcast(ref, src=0)
    same_ranks = torch.equal(ref, out["p2p"][0])
    flags = torch.tensor([int(same_ranks and same_traj)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        full = DataShardedHMC(model, theta0, torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), step=step, num_steps=L,
                              seed=3, exchange="local")
    dist.barrier()
    if rank == 0:
        fs, ft, fa = full.run(num_epochs=T, num_burnin_epochs=0)
        torch.cuda.synchronize()
        rel = ((fs - out["p2p"][0]).abs().max() / fs.abs().max()).item()
        acc_same = torch.equal(fa, out["p2p"][2])
        print(f"persistent == per-evaluation launches bitwise: {same_traj}; p2p == nccl bitwise: {same} (max abs diff {err_modes:.2e}); "
              f"identical on all ranks: {bool(flags.item())}; "
              f"vs one shard: rel {rel:.2e}, accepts equal {acc_same}; acceptance {out['p2p'][2].float().mean().item():.2f}", flush=True)
        ok = (err_modes < 1e-6) and bool(flags.item()) and rel < 1e-5 and acc_same
        print("CHECK", "OK" if ok else "FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
