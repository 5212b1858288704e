"""world_size-2 gloo tests (CPU) of the two multi-GPU decompositions of the path:
 (i) chain sharding -- no data-path collective: a rank running chains [lo, hi) with chain_offset = lo reproduces exactly
     those chains of the single-rank run (device code executed through tests/hostsim);
 (ii) data sharding -- one exchange step per evaluation: the all-reduced partial (loglik, gradient) sums over row shards
     equal the unsharded evaluation (the oracle stands in for the per-rank kernel on CPU);
 (iii) the rank-ordered inbox summation that replaces the all-reduce on the GPUs (csrc/datapar.cu: dp_post_kernel)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from eeyore_b200._native import RunParams, F64
from eeyore_b200.samplers.data_sharded_hmc import shard_rows
from helpers import data_of, spec_of


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _hmc_hostsim(lib, theta0, x, y, chain_offset, T, L, step, seed):
    Cn, n = theta0.shape
    loc, scale = np.zeros(n), np.full(n, 3 ** 0.5)
    theta = theta0.copy()
    lt = np.empty(Cn); g = np.empty_like(theta)
    lib.hostsim_eval(2321, F64, Cn, theta.ctypes.data, x.ctypes.data, y.ctypes.data, x.shape[0], loc.ctypes.data,
                     scale.ctypes.data, 0, 0.0, lt.ctypes.data, g.ctypes.data)
    out = np.zeros((T, Cn, n))
    p = RunParams()
    p.n_chains, p.n_iters, p.n_burnin, p.thin = Cn, T, 0, 1
    p.step, p.num_steps, p.rng_mode, p.seed, p.chain_offset = step, L, 0, seed, chain_offset
    p.x, p.y, p.n_rows = x.ctypes.data, y.ctypes.data, x.shape[0]
    p.prior_loc, p.prior_scale = loc.ctypes.data, scale.ctypes.data
    p.theta, p.target, p.grad = theta.ctypes.data, lt.ctypes.data, g.ctypes.data
    p.out_samples, p.ss_iter, p.ss_chain, p.ss_param = out.ctypes.data, Cn * n, n, 1
    assert lib.hostsim_run(2, 2321, F64, C.byref(p)) == 0
    return out


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import __graft_entry__ as g
        lib = C.CDLL(str(g.build_hostsim()))
        lib.hostsim_eval.argtypes = [C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 2 + \
                                    [C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        lib.hostsim_run.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(RunParams)]
        # (i) chain sharding
        Cn, T, L, step, seed = 10, 6, 4, 0.4, 17
        x, y = data_of("2321", np.float64)
        x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
        theta0 = np.random.default_rng(0).normal(size=(Cn, 20))
        per = Cn // world
        lo, hi = rank * per, (rank + 1) * per
        mine = torch.from_numpy(_hmc_hostsim(lib, theta0[lo:hi], x, y, lo, T, L, step, seed))
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)                       # only the final gather communicates
        if rank == 0:
            full = _hmc_hostsim(lib, theta0, x, y, 0, T, L, step, seed)
            assert np.array_equal(torch.cat(gathered, dim=1).numpy(), full)
        # (ii) data sharding
        n = 1003
        rng = np.random.default_rng(1)
        xs = rng.normal(size=(n, 16)); ys = (rng.uniform(size=(n, 1)) < 0.5).astype(np.float64)
        spec = oracle.MLPSpec([16, 64, 64, 1], loss=oracle.mlp.BINARY)
        th = rng.normal(size=(1, spec.num_params)) * 0.2
        lo, hi = shard_rows(n, world, rank)
        h = oracle.mlp.forward(spec, th, xs[lo:hi])
        ll, seed_ = oracle.mlp._loss_seed(spec, h[-1], ys[lo:hi])
        zero_prior = (np.zeros(spec.num_params), np.full(spec.num_params, 1e150))
        _, gl = oracle.log_target_grad(spec, th, xs[lo:hi], ys[lo:hi], *zero_prior)
        sums = torch.from_numpy(np.concatenate([ll, gl[0]]))
        dist.all_reduce(sums)                                  # the path's one exchange step: 1 + P doubles
        ll_full = oracle.log_lik(spec, th, xs, ys)[0]
        _, g_full = oracle.log_target_grad(spec, th, xs, ys, *zero_prior)
        assert abs(sums[0].item() - ll_full) < 1e-9 * abs(ll_full)
        assert np.max(np.abs(sums[1:].numpy() - g_full[0])) < 1e-9 * np.max(np.abs(g_full[0]))
        # (iii) the peer-store exchange of dp_post_kernel, emulated: every rank's sums land in slot [src] of every rank's inbox
        # (an all-gather stands in for the NVLink stores) and each rank adds the slots in rank order -> bit-identical totals
        mine = torch.from_numpy(np.concatenate([ll, gl[0]]))
        inbox = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(inbox, mine)
        total = torch.zeros_like(mine)
        for src in range(world):
            total += inbox[src]
        both = [torch.empty_like(total) for _ in range(world)]
        dist.all_gather(both, total)
        assert all(torch.equal(both[0], b) for b in both)              # replicated chain state stays in lock step
        assert torch.allclose(total, sums, rtol=1e-13, atol=0)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_shard_rows_partition():
    for n in (1, 5, 127, 1000, 8388608):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(lo % 4 == 0 for lo, hi in spans if hi > lo)
