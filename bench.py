#!/usr/bin/env python
"""Benchmark of the sampler hot path (BASELINE.json metric: log-target+gradient evaluations per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg2|cfg3|cfg5] [--impl ours|reference] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

The default line is config 4 (below); its `extra` object carries, measured in the same process at the same N, every other
BASELINE configuration: cfg4_strong (524,288 chains in total, sharded), cfg4_thin1 (every iteration saved), cfg4_diagnostics
(1,000 saved iterations per chain + on-device multi-ESS / ACF of every chain), cfg2, cfg3, cfg4_f32 / cfg2_f32 (the fp32 chain
kernels), cfg1 (one chain, wall time) and cfg5 (the data-sharded path with the peer-store exchange).

Workload (default cfg4 = BASELINE.json configs[3], the configuration the 1/2/4/8-GPU metric is quoted on):
MLP 2-3-2-1 on XOR, HMC with 10 leapfrog steps, 524,288 independent fp64 chains PER GPU (chains are sharded across
ranks with no data-path collective => weak scaling), on-device Philox noise, random-init chain states.
One step = one fused sampler launch advancing every chain by ITERS HMC iterations.
Evaluations are counted as chains x leapfrog steps (the metric's own definition); the reference executes one
extra, redundant, gradient evaluation per iteration (eeyore/samplers/hmc.py:104) which is not counted on either arm.

Prints ONE JSON line (rank 0).  `value` is measured with the chain state resident in HBM; `e2e` is measured through
the public sampler API with HOST (pinned) buffers: per step the chain states are copied host->device, the sampler
runs, and the final states / targets / accept counts are copied device->host.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

S3 = 3.0 ** 0.5

WORKLOADS = {
    # BASELINE.json configs[3]
    "cfg4": dict(name="cfg4: MLP 2-3-2-1 XOR, HMC L=10, 524288 chains/GPU, fp64", dims=[2, 3, 2, 1], data="xor",
                 loss="binary_classification", chains=524288, step=0.3, num_steps=10, iters=50, thin=10,
                 flops_per_eval=368, dtype="f64"),
    # BASELINE.json configs[2] (SMMALA is builder-defined: absent from the reference snapshot)
    "cfg3": dict(name="cfg3: MLP 2-3-2-1 noisy-XOR-shaped N=200, SMMALA (Fisher metric, in-warp Cholesky), 16384 chains/GPU, fp64",
                 dims=[2, 3, 2, 1], data="noisy_xor", loss="binary_classification", chains=16384, step=0.02, num_steps=1,
                 iters=20, thin=5, flops_per_eval=14480 + 160000 + 2667 + 800, dtype="f64", kind="smmala"),
    # BASELINE.json configs[4]: one chain, data sharded over the ranks, NCCL all-reduce per evaluation (strong scaling)
    "cfg5": dict(name="cfg5: MLP 16-64-64-1, 8388608 synthetic rows sharded over the GPUs, HMC L=10, fp32",
                 dims=[16, 64, 64, 1], data="teacher", loss="binary_classification", chains=1, step=4e-5, num_steps=10,
                 iters=2, thin=1, flops_per_eval=29056 * 8388608, rows=8388608, dtype="f32", kind="datapar"),
    # BASELINE.json configs[1]
    "cfg2": dict(name="cfg2: MLP 4-3-3 iris-shaped N=150, HMC L=10, 4096 chains/GPU, fp64", dims=[4, 3, 3], data="iris",
                 loss="multiclass_classification", chains=4096, step=0.15, num_steps=10, iters=20, thin=5,
                 flops_per_eval=15408, dtype="f64"),
}


def synthetic_data(w):
    """cfg4: the XOR truth table; cfg2: iris-shaped synthetic data (3 Gaussian classes, 50 rows each, seed 1)."""
    if w["data"] == "xor":
        x = np.array([[0, 0], [0, 1], [1, 0], [1, 1]], dtype=np.float64)
        y = np.array([[0], [1], [1], [0]], dtype=np.float64)
        return x, y
    if w["data"] == "noisy_xor":          # 50 points per XOR corner + N(0, 0.15^2), seed 3
        rng = np.random.default_rng(3)
        corners = np.array([[0, 0], [0, 1], [1, 0], [1, 1]], dtype=np.float64)
        x = np.concatenate([c + 0.15 * rng.normal(size=(50, 2)) for c in corners])
        y = np.concatenate([np.full((50, 1), float(int(c[0]) ^ int(c[1]))) for c in corners])
        return x, y
    rng = np.random.default_rng(1)
    centres = rng.normal(size=(3, 4)) * 2.0
    x = np.concatenate([centres[k] + 0.5 * rng.normal(size=(50, 4)) for k in range(3)])
    y = np.eye(3)[np.repeat(np.arange(3), 50)]
    return x, y


# ------------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle-reason poller for the timed region (rank 0 only), through NVML in-process (the data source of
    `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*`).  Spawning nvidia-smi itself next to a
    launch-heavy timed region is avoided on purpose: its start-up / teardown (NVML init and shutdown) was measured to
    stall kernel submission for hundreds of milliseconds on these boxes.  NVML is initialised once, before the warm-up;
    only samples taken between mark_begin() and mark_end() are summarised."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index, enabled=True):
        self.index, self.rows, self.enabled, self.h = index, [], enabled, None
        self.t0, self.t1, self._stop = 0.0, float("inf"), False

    def start(self):
        if not self.enabled:
            return self
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.h = None
        return self

    def _poll(self):
        while not self._stop:
            try:
                sm = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                why = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.perf_counter(), sm, why))
            except Exception:
                pass
            time.sleep(0.05)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self._stop = True

    def summary(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "unavailable"}
        rows = [r for r in self.rows if self.t0 <= r[0] <= self.t1]
        if not rows:
            rows = self.rows[-3:]
        sm = [r[1] for r in rows]
        mask = 0
        for r in rows:
            mask |= r[2]
        reasons = sorted(n for n, b in self.REASONS.items() if mask & b)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.max_sm), "reasons": reasons,
                "samples": len(sm), "source": "NVML (nvmlDeviceGetClockInfo / CurrentClocksThrottleReasons), 50 ms period"}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (numpy restatement of the reference path), all host cores
# ------------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    wname, chains, iters, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import oracle
    from oracle.mlp import MLPSpec
    w = WORKLOADS[wname]
    spec = MLPSpec(w["dims"], loss=w["loss"])
    x, y = synthetic_data(w)
    p = spec.num_params
    rng = np.random.default_rng(seed)
    theta = rng.normal(size=(chains, p)) * (1.0 if w["data"] == "xor" else 0.3)
    z, u = rng.normal(size=(iters, chains, p)), rng.uniform(size=(iters, chains))
    t0 = time.perf_counter()
    if w.get("kind") == "smmala":
        oracle.smmala_run(spec, x, y, np.zeros(p), np.full(p, S3), theta * 0.5, z, u, w["step"])
    else:
        oracle.hmc_run(spec, x, y, np.zeros(p), np.full(p, S3), theta, z, u, w["step"], w["num_steps"])
    return time.perf_counter() - t0


def _cpu_worker_dp(args):
    """cfg5 on the CPU: one log-target + gradient evaluation of the numpy oracle port over a slice of `rows` synthetic rows
    (the rows are independent, so the host cores shard them exactly as the GPUs do)."""
    rows, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import oracle
    from oracle.mlp import MLPSpec
    w = WORKLOADS["cfg5"]
    spec = MLPSpec(w["dims"], loss=w["loss"])
    p = spec.num_params
    rng = np.random.default_rng(seed)
    theta = rng.normal(size=(1, p)) * 0.1
    t_total, chunk = 0.0, 65536
    for c0 in range(0, rows, chunk):
        n = min(chunk, rows - c0)
        x = rng.normal(size=(n, 16))
        y = (rng.uniform(size=(n, 1)) < 0.5).astype(np.float64)
        t0 = time.perf_counter()
        oracle.log_target_grad(spec, theta, x, y, np.zeros(p), np.full(p, S3))
        t_total += time.perf_counter() - t0
    return t_total


def cpu_datapar_throughput(pool, procs, sample_rows, n_total):
    """Full-data-set evaluations per second of the oracle port: `sample_rows` rows split over the host cores, scaled to
    n_total rows (the cost is linear in the row count)."""
    per = max(1, sample_rows // procs)
    t0 = time.perf_counter()
    pool.pool.map(_cpu_worker_dp, [(per, s) for s in range(procs)], chunksize=1)
    wall = time.perf_counter() - t0
    return (per * procs / n_total) / wall, wall


class CpuPool:
    """One worker process per host core, started once (outside any timed region)."""

    def __init__(self, procs):
        import multiprocessing as mp
        self.procs = procs
        self.pool = mp.get_context("spawn").Pool(procs)
        self.pool.map(_cpu_worker, [("cfg4", 8, 1, s) for s in range(procs)])   # import / warm every worker

    def throughput(self, wname, chains_per_proc, iters):
        """evals/s of the oracle port with every worker running chains_per_proc chains for `iters` HMC iterations."""
        w = WORKLOADS[wname]
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, [(wname, chains_per_proc, iters, s) for s in range(self.procs)], chunksize=1)
        wall = time.perf_counter() - t0
        return self.procs * chains_per_proc * iters * w["num_steps"] / wall, wall

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_sizes(wname):
    w = WORKLOADS[wname]
    if wname == "cfg4":
        return 4096, 3          # chains per process, iterations
    if wname == "cfg3":
        return 64, 2
    return 128, 2


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    procs = host_cores()
    if w.get("kind") == "datapar":
        n_total = args.rows or w["rows"]
        sample_rows = 1 << 20
        pool = CpuPool(procs)
        for _ in range(args.warmup):
            cpu_datapar_throughput(pool, procs, 1 << 16, n_total)
        walls, vals = [], []
        for _ in range(args.steps):
            v, wall = cpu_datapar_throughput(pool, procs, sample_rows, n_total)
            vals.append(v)
            walls.append(wall)
        pool.close()
        value = len(vals) / sum(1.0 / v for v in vals)
        sample = (f"each step: one evaluation over {sample_rows} of the {n_total} rows, split over {procs} processes (one per host "
                  f"core) of the numpy oracle port (oracle/mlp.py); evaluations/s scaled linearly to {n_total} rows")
        print(json.dumps({
            "impl": "reference", "metric": "log_target_grad_evals_per_sec", "value": value, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "rows_total": n_total, "evals_counted_per_iteration": w["num_steps"]},
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return
    cpp, iters = cpu_sample_sizes(args.workload)
    pool = CpuPool(procs)
    for _ in range(args.warmup):
        pool.throughput(args.workload, max(cpp // 8, 16), 1)
    t_steps, evals = [], 0
    for _ in range(args.steps):
        v, wall = pool.throughput(args.workload, cpp, iters)
        t_steps.append(wall)
        evals += procs * cpp * iters * (1 if w.get("kind") == "smmala" else w["num_steps"])
    pool.close()
    total = sum(t_steps)
    value = evals / total
    sample = (f"each step: {procs} processes (one per host core) x {cpp} chains x {iters} HMC iterations "
              f"(L={w['num_steps']}) of the numpy oracle port (oracle/samplers.py), chain-batched")
    print(json.dumps({
        "impl": "reference", "metric": "log_target_grad_evals_per_sec", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
        "config": {"workload": w["name"], "evals_counted_per_iteration": w["num_steps"]},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(index):
    """Pins this rank's threads to the CPUs NVML lists as local to its GPU (before any pinned host buffer is allocated, so the
    pages land on that socket): with eight ranks streaming 36 GB/s each into host memory the traffic should not cross the
    socket interconnect.  Best effort: a container may not allow it."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%d cpus local to GPU %d" % (len(cpus), phys)
    except Exception as e:      # noqa: BLE001
        return "not bound (%s)" % type(e).__name__
    return "not bound"


class Ctx:
    """Process-wide state of one bench run: rank / device / process group and the timing helpers every record uses."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa = bind_to_gpu_numa_node(self.local) if self.world > 1 else None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.args = args
        self.flush_buf = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    def flush_l2(self, k=0):
        """Overwrites 160 MB of device memory (the L2 is 126 MB)."""
        if self.flush_buf is None:
            self.flush_buf = self.torch.empty(160 << 20, dtype=self.torch.uint8, device=self.dev)
        self.flush_buf.fill_(k & 0xFF)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def build_model(ctx, w):
    torch = ctx.torch
    from torch.distributions import Normal
    from torch.utils.data import DataLoader
    from eeyore_b200.constants import loss_functions
    from eeyore_b200.datasets import XYDataset
    from eeyore_b200.models.mlp import MLP, Hyperparameters
    dt = torch.float64 if w["dtype"] == "f64" else torch.float32
    x, y = synthetic_data(w)
    ds = XYDataset(torch.from_numpy(x).to(dt), torch.from_numpy(y).to(dt))
    nl = len(w["dims"]) - 1
    binary = w["loss"] == "binary_classification"
    hp = Hyperparameters(w["dims"], nl * [True], (nl - 1) * [torch.sigmoid] + [torch.sigmoid if binary else None])
    model = MLP(loss=loss_functions[w["loss"]], hparams=hp, dtype=dt, device=ctx.dev)
    P = model.num_params()
    model.prior = Normal(torch.zeros(P, dtype=dt), S3 * torch.ones(P, dtype=dt))
    return model, DataLoader(ds, batch_size=len(ds)), dt, x, y


def make_sampler(w, model, loader, theta0, seed, thin, lanes=0, chain=None):
    if w.get("kind") == "smmala":
        from eeyore_b200.samplers import SMMALA
        return SMMALA(model, theta0=theta0, dataloader=loader, step=w["step"], seed=seed, thin=thin, chain=chain)
    from eeyore_b200.samplers import HMC
    return HMC(model, theta0=theta0, dataloader=loader, step=w["step"], num_steps=w["num_steps"], seed=seed, thin=thin,
               lanes_per_chain=lanes, chain=chain)


def bench_chains(ctx, wname, steps, warmup, chains=None, thin=None, iters=None, e2e=True, clocks=None, scaling="weak",
                 dtype=None):
    """One chain-sharded workload: `value` with the chain state resident in HBM (one fused launch per step) and `e2e`
    through the public sampler API with pinned HOST buffers: per step the chain states are copied host->device, the sampler
    runs, and everything a ChainList would hold afterwards -- the saved samples, their targets and accept flags, the final
    states and accept counts -- is copied device->host."""
    torch = ctx.torch
    args = ctx.args
    w = dict(WORKLOADS[wname])
    if dtype:
        w["dtype"] = dtype
    model, loader, dt, x, y = build_model(ctx, w)
    P = model.num_params()
    C = chains or args.chains or w["chains"]
    iters = iters or args.iters or w["iters"]
    thin = thin or w["thin"]
    L = w["num_steps"]
    kind = w.get("kind", "hmc")
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    gen = torch.Generator().manual_seed(1000 + rank)
    theta_host = (torch.randn(C, P, generator=gen, dtype=dt) * {"xor": 1.0, "noisy_xor": 0.5}.get(w["data"], 0.3)).pin_memory()

    sampler = make_sampler(w, model, loader, theta_host.to(dev), 12345, thin, args.lanes)
    sampler.chain_offset = rank * C          # global chain ids => results independent of the sharding

    def step_resident():
        sampler._device_blocks = []          # saved states of the previous step are dropped (buffer is recycled)
        sampler.counter.reset()
        sampler.run(num_epochs=iters, num_burnin_epochs=0)

    for _ in range(warmup):
        step_resident()
    ctx.barrier()
    esz = theta_host.element_size()
    n_saved = (iters + thin - 1) // thin
    hbm_bytes = C * (2 * (2 * P + 1) * esz + n_saved * (P * esz + esz + 1) + 4)
    flush = hbm_bytes <= 126e6     # working set below the L2 size: flush between timed steps (outside the event pairs)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ctx.barrier()
    if clocks:
        clocks.mark_begin()
    for k in range(steps):
        if flush:
            ctx.flush_l2(k)
        ev0[k].record()
        step_resident()
        ev1[k].record()
    ctx.barrier()
    if clocks:
        clocks.mark_end()
    per_launch_ms = [ev0[k].elapsed_time(ev1[k]) for k in range(steps)]
    t_local = (sum(per_launch_ms) if flush else ev0[0].elapsed_time(ev1[-1])) * 1e-3
    t_res = ctx.max_over_ranks(t_local)
    evals_step = C * iters * (1 if kind == "smmala" else L)
    value = world * evals_step * steps / t_res
    acc_rate = sampler.get_chain().acceptance().mean().item()
    rec = {
        "value": value, "unit": "evals/s", "ms_per_step": 1e3 * t_res / steps, "scaling": scaling, "dtype": w["dtype"],
        "config": {"workload": w["name"], "chains_per_gpu": C, "chains_total": C * world, "iterations_per_step": iters,
                   "num_steps": L, "step_size": w["step"], "thin": thin, "evals_counted_per_iteration": 1 if kind == "smmala" else L,
                   "acceptance_rate": acc_rate,
                   "l2": ("chain state + saved samples per launch (%.0f MB) exceed the 126 MB L2" % (hbm_bytes / 1e6)) if not flush else
                         ("chain state + saved samples per launch are %.0f MB (below the 126 MB L2): 160 MB of device memory are "
                          "overwritten between timed steps to flush it" % (hbm_bytes / 1e6))},
        "gpu_launches": steps,
        "_per_launch_ms": per_launch_ms, "_evals_step": evals_step, "_hbm_bytes": hbm_bytes, "_P": P, "_C": C, "_dt": dt,
    }
    del sampler

    if e2e:
        # The chains are independent, so the step is cut into `nb` chain batches, each with its own persistent sampler and
        # stream: batch b + 1's host->device copy runs under batch b's kernel.  The samplers run with host_output = True: the
        # kernel stores every saved state (sample, target, accept flag) straight into pinned host memory -- the stores ARE the
        # device->host transfer (posted PCIe writes under the computation, measured at +0.2 ms on the 15 ms kernel), so no
        # staging copy queues up behind the kernels; only the final states / targets / accept counts are copied afterwards.
        # Philox is keyed by the global chain id, so the results do not depend on the batching.  reset(theta_host) re-uses
        # every buffer of the sampler (no construction or allocation inside the timed region).
        nb = max(1, min(args.e2e_batches, C // 32768))
        bounds = [(b * C // nb, (b + 1) * C // nb) for b in range(nb)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(nb)]
        samplers = []
        for (lo, hi) in bounds:
            s = make_sampler(w, model, loader, theta_host[lo:hi], 999, thin, args.lanes)
            s.chain_offset = rank * C + lo
            s.host_output = True
            if kind == "smmala":
                # warp-per-chain kernel: lane = parameter.  Chain-major blocks make a chain's saved sample one contiguous 160-byte
                # store; in the chain-minor layout every lane's 8 bytes would cross PCIe as a transaction of their own
                s.sample_layout = "cnp"
            samplers.append(s)
        pin = lambda *shape, dtype=dt: torch.empty(*shape, dtype=dtype).pin_memory()
        out_theta, out_lt, out_acc = pin(C, P), pin(C), pin(C, dtype=torch.int32)

        def step_e2e():
            cur = torch.cuda.current_stream()
            for b, ((lo, hi), st, s) in enumerate(zip(bounds, streams, samplers)):
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    s.reset(theta_host[lo:hi])
                    s.run(num_epochs=iters, num_burnin_epochs=0)
                    if s.host_current is None:       # samplers without the final-state host outputs (SMMALA): staged copies
                        out_theta[lo:hi].copy_(s.current["sample"], non_blocking=True)
                        out_lt[lo:hi].copy_(s.current["target_val"], non_blocking=True)
                        out_acc[lo:hi].copy_(s.acceptance_counts(), non_blocking=True)
            for st in streams:
                cur.wait_stream(st)
            cur.synchronize()
            blk = samplers[-1]._device_blocks[-1]              # pinned host tensors, complete after the synchronisation
            fin = samplers[-1].host_current
            last_lt = out_lt[-1].item() if fin is None else float(fin["target_val"][-1]) + float(fin["sample"][-1, 0])
            return last_lt + float(blk["sample"][-1, 0, -1]) + float(blk["target_val"][-1, -1])

        for _ in range(max(2, warmup // 2)):
            step_e2e()
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            step_e2e()
        e1.record()
        ctx.barrier()
        t_e2e = ctx.max_over_ranks(e0.elapsed_time(e1) * 1e-3)
        h2d = theta_host.numel() * esz
        host_blocks = [t for s in samplers for t in s._device_blocks[-1].values()]
        assert all(t.device.type == "cpu" for t in host_blocks)
        finals = [t for s in samplers if s.host_current is not None for t in s.host_current.values()]
        if not finals:
            finals = [out_theta, out_lt, out_acc]
        d2h = sum(t.numel() * t.element_size() for t in (*finals, *host_blocks))
        rec["e2e"] = {"value": world * evals_step * steps / t_e2e, "unit": "evals/s", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e / steps, "chain_batches": nb,
                      "returns": "every saved sample [n_saved, P, C], its target and accept flag, the final states / targets / "
                                 "accept counts: what the ChainList objects of the run hold",
                      "note": "public sampler API (reset + run, host_output) on %d persistent per-batch samplers / streams; chain "
                              "states arrive by pinned host->device copies; saved states and (HMC) the final states / targets / "
                              "accept counts leave as the kernel's own stores into pinned host memory" % nb}
        if ctx.numa:
            rec["e2e"]["host_binding"] = ctx.numa
        rec["gpu_launches_e2e"] = (4 * nb) * steps   # per batch: eval + two transposes of reset, the fused run
        del samplers
    torch.cuda.empty_cache()
    return rec


def chain_roofline(ctx, rec, w):
    """FMA roofline of the dominant kernel of a chain workload (sampler_kernel / smmala_kernel): algorithmic FLOPs per launch /
    average launch time against the live-measured FMA peak of this device."""
    import ctypes
    from eeyore_b200 import _native as nv
    dt = rec["_dt"]
    peak = ctypes.c_double()
    nv.check(nv.lib().eeyore_b200_fma_peak(nv.DTYPE_IDS[dt], 2000, ctypes.byref(peak)))
    avg_launch_s = float(np.mean(rec["_per_launch_ms"])) * 1e-3
    achieved = w["flops_per_eval"] * rec["_evals_step"] / avg_launch_s / 1e12
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    hbm_peak = json.loads(peaks_file.read_text())["hbm_gbs"] if peaks_file.exists() else 6650.0
    out = {"bound": "fp64_fma" if dt == ctx.torch.float64 else "fp32_fma", "achieved": achieved, "peak": peak.value,
           "unit": "TFLOP/s", "frac": achieved / peak.value, "traffic": None,
           "peak_source": "measured live by eeyore_b200_fma_peak (dependent-free FMA chains, this device)",
           "algorithmic_flops_per_eval": w["flops_per_eval"], "avg_launch_ms": avg_launch_s * 1e3,
           "hbm_view": {"algorithmic_bytes_per_launch": rec["_hbm_bytes"], "achieved_gbs": rec["_hbm_bytes"] / avg_launch_s / 1e9,
                        "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks_file.exists() else "fallback"}}
    return out


def strip(rec):
    return {k: v for k, v in rec.items() if not k.startswith("_")}


def bench_diagnostics(ctx, reps=2):
    """BASELINE config 4's second half: HMC sampling followed by on-device multi-ESS and autocorrelation of EVERY chain.
    The saved-sample ring of all chains at once would be 84 GB (524,288 x 1,000 x 20 fp64), so the chains go through in
    chunks: sample a chunk for n iterations (thin = 1, every chain's samples contiguous: sample_layout 'cnp'), run the
    diagnostics kernel on it, keep [ESS, ACF] (C x (1 + 11 x 20) values), drop the ring.  Reported: the two stage times."""
    torch = ctx.torch
    from eeyore_b200 import stats as st
    from eeyore_b200.chains import ChainList
    args = ctx.args
    w = WORKLOADS["cfg4"]
    model, loader, dt, x, y = build_model(ctx, w)
    P = model.num_params()
    C = args.chains or w["chains"]
    n, max_lag, chunk = args.diag_samples, 10, min(args.diag_chunk, C)
    n_chunks = (C + chunk - 1) // chunk
    gen = torch.Generator().manual_seed(2000 + ctx.rank)
    theta_host = torch.randn(C, P, generator=gen, dtype=dt).pin_memory()
    ess = torch.empty(C, dtype=dt, device=ctx.dev)
    acf = torch.empty(C, max_lag + 1, P, dtype=dt, device=ctx.dev)
    status = torch.empty(C, dtype=torch.int32, device=ctx.dev)
    samplers = []
    for k in range(n_chunks):
        lo, hi = k * chunk, min(C, (k + 1) * chunk)
        s = make_sampler(w, model, loader, theta_host[lo:hi], 4242, 1, args.lanes, chain=ChainList(keys=["sample", "accepted"]))
        s.chain_offset = ctx.rank * C + lo
        s.sample_layout = "cnp"
        samplers.append(s)

    left = []

    def one_pass(timed):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * n_chunks + 1)]
        ev[0].record()
        for k, s in enumerate(samplers):
            lo, hi = k * chunk, min(C, (k + 1) * chunk)
            s._device_blocks = []
            s.counter.reset()
            s.run(num_epochs=n, num_burnin_epochs=0)
            ev[2 * k + 1].record()
            ring = s._device_blocks[-1]["sample"]                   # [n, P, c] view of the chain-major [c, n, P] buffer
            # chains whose INSE estimate is still not positive definite after 15 lag pairs (a third of a per cent: antithetic
            # HMC chains, which end as 'Not enough samples' after n / 2 lag pairs) are left undecided here, their samples kept
            # (160 KB each), and all of them are finished side by side in ONE launch after the last chunk
            out = st.chain_stats(ring, layout="npc", want=("ess",), max_lag=max_lag, check=False, defer="leave")
            ess[lo:hi], acf[lo:hi], status[lo:hi] = out["ess"], out["acf"], out["status"]
            todo = (out["status"] == 3).nonzero().flatten()
            if todo.numel():
                left.append((todo + lo, ring.permute(2, 0, 1)[todo]))
            s._device_blocks = []
            del ring, out
            ev[2 * k + 2].record()
        e_tail0, e_tail1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_tail0.record()
        n_left = 0
        if left:
            idx = torch.cat([i for i, _ in left])
            n_left = int(idx.numel())
            out = st.chain_stats(torch.cat([x_ for _, x_ in left]), layout="cnp", want=("ess",), max_lag=max_lag, check=False,
                                 defer=False)
            ess[idx], acf[idx], status[idx] = out["ess"], out["acf"], out["status"]
            left.clear()
        e_tail1.record()
        torch.cuda.synchronize()
        t_sample = sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(n_chunks)) * 1e-3
        t_stats = (sum(ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(n_chunks)) + e_tail0.elapsed_time(e_tail1)) * 1e-3
        return t_sample, t_stats, n_left, e_tail0.elapsed_time(e_tail1) * 1e-3

    one_pass(False)
    ctx.barrier()
    ts, tt, nl, tl = zip(*[one_pass(True) for _ in range(reps)])
    ctx.barrier()
    t_sample, t_stats = ctx.max_over_ranks(float(np.mean(ts))), ctx.max_over_ranks(float(np.mean(tt)))
    ok = status == 0
    # algorithmic work of the diagnostics of one chain: mean (n P), lags 0 and 1 (2 n P^2 each), one pass per further lag pair
    # (2 n P^2, pair-summed), ACF (2 n P (max_lag + 1)); bytes: one read of the chain per pass
    summary = {
        "value": ctx.world * C / (t_sample + t_stats), "unit": "chains/s (sampled for %d iterations and diagnosed)" % n,
        "chains_per_gpu": C, "chains_total": C * ctx.world, "samples_per_chain": n, "chunk_chains": chunk, "acf_max_lag": max_lag,
        "sampling_s": t_sample, "diagnostics_s": t_stats, "diagnostics_over_sampling": t_stats / t_sample,
        "sampling_evals_per_s": ctx.world * C * n * w["num_steps"] / t_sample,
        "diagnosed_chains_per_s": ctx.world * C / t_stats,
        "ring_bytes_per_chunk": chunk * n * P * 8,
        "ess_mean": float(ess[ok].mean()) if bool(ok.any()) else None, "ess_min": float(ess[ok].min()) if bool(ok.any()) else None,
        "chains_not_enough_samples": int((~ok).sum()), "chains_finished_in_the_last_launch": int(nl[-1]),
        "last_launch_s": float(np.mean(tl)),
        "acf_lag1_mean": float(acf[:, 1].mean()), "acf_lag10_mean": float(acf[:, max_lag].mean()),
        "kernel": "chain_stats_kernel<double, 5> (warp per chain, TMA-fed ring, register Cholesky / LU)",
    }
    del samplers, ess, acf
    torch.cuda.empty_cache()
    return summary


def bench_cfg1(ctx):
    """BASELINE configs[0] through the public API: MLP 2-2-1 on XOR, ONE MALA chain, 1,100 iterations (110 burn-in), fp64 --
    one fused launch.  Wall time of sampler.run() including the synchronisation, best of three (the reference: 1.95 - 2.3 s)."""
    torch = ctx.torch
    from eeyore_b200.samplers import MALA
    w = dict(name="cfg1", dims=[2, 2, 1], data="xor", loss="binary_classification", dtype="f64")
    model, loader, dt, x, y = build_model(ctx, w)
    theta0 = model.prior.sample()
    best = None
    for rep in range(4):
        s = MALA(model, theta0=theta0, dataloader=loader, step=1.74, seed=rep)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.run(num_epochs=1100, num_burnin_epochs=110)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        if rep > 0:
            best = t if best is None else min(best, t)
    ch = s.get_chain()
    return {"wall_ms_run_1100_iterations": 1e3 * best, "value": 1100 / best, "unit": "evals/s (one chain: latency, not throughput)",
            "saved_samples": len(ch), "acceptance_rate": ch.acceptance_rate(), "launches": 1,
            "multi_ess_of_the_chain": ch.multi_ess()}


def run_ours(args):
    ctx = Ctx(args)
    torch = ctx.torch
    clocks = ClockSampler(ctx.local, enabled=(ctx.rank == 0 and not os.environ.get("EEYORE_BENCH_NO_CLOCKS"))).start()
    wname = args.workload
    w = WORKLOADS[wname]
    if w.get("kind") == "datapar":
        line = bench_datapar(ctx, args.steps, args.warmup, clocks=clocks, full=True)
    else:
        head = bench_chains(ctx, wname, args.steps, args.warmup, clocks=clocks, dtype=args.dtype)
        clocks.stop()
        line = None
        if ctx.rank == 0:
            roof = chain_roofline(ctx, head, w)
            tfile = ROOT / "profiles" / "r02_traffic_cfg4.json"
            if wname == "cfg4" and not args.chains and not args.iters and not args.dtype and tfile.exists():
                rec = json.loads(tfile.read_text())
                roof["traffic"] = rec["dram_bytes_total"]   # dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full
                # instruction-level view: the FP64 pipe takes one warp instruction every two cycles per SM sub-partition
                sm_mhz = (clocks.summary().get("sm_mhz") or 1965.0)
                sms = torch.cuda.get_device_properties(ctx.dev).multi_processor_count
                pipe_rate = sms * 2 * sm_mhz * 1e6
                n64 = rec["fp64_pipe_warp_instructions_per_warp_evaluation"]
                roof["fp64_pipe_view"] = {
                    "fp64_pipe_warp_instructions_per_warp_evaluation": n64,
                    "warp_instructions_per_warp_evaluation": rec["warp_instructions_per_warp_evaluation"],
                    "pipe_rate_warp_instructions_per_s": pipe_rate,
                    "frac": (head["value"] / ctx.world / 32.0) * n64 / pipe_rate,
                    "ncu_fp64_pipe_cycles_active_pct": rec["fp64_pipe_cycles_active_pct"],
                    "source": "instruction counts from the committed ncu capture (profiles/r02_traffic_cfg4.json); rate = SMs x 2 per cycle x SM clock sampled during the timed region"}
            line = {"metric": "log_target_grad_evals_per_sec", "value": head["value"], "unit": "evals/s", "n_gpus": ctx.world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic",
                    "config": dict(head["config"], evals_reference_executes_per_iteration=w["num_steps"] + 1,
                                   rng="philox4x32-10 on device"),
                    "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "gpu_launches_e2e": head["gpu_launches_e2e"],
                    "clocks": clocks.summary(), "roofline": roof}
    # ---- every other BASELINE configuration under the same clock (sub-records; the headline stays the line's own keys) ----
    extra = {}
    if args.extras and wname == "cfg4" and not args.dtype:
        k = max(3, args.steps // 2)
        wu = max(3, args.warmup // 2)
        if ctx.world > 1:   # config 4 as BASELINE states it: 524,288 chains IN TOTAL, sharded over the GPUs
            r = bench_chains(ctx, "cfg4", k, wu, chains=WORKLOADS["cfg4"]["chains"] // ctx.world, e2e=False, scaling="strong")
            extra["cfg4_strong"] = strip(r)
        else:
            extra["cfg4_strong"] = {"note": "one GPU: identical to the headline (524,288 chains on this GPU)", "value": None}
        r = bench_chains(ctx, "cfg4", k, wu, thin=1, e2e=False)
        extra["cfg4_thin1"] = strip(r)
        extra["cfg4_diagnostics"] = bench_diagnostics(ctx)
        for name in ("cfg2", "cfg3"):
            r = bench_chains(ctx, name, k, wu)
            if ctx.rank == 0:
                r["roofline"] = chain_roofline(ctx, r, WORKLOADS[name])
            extra[name] = strip(r)
        for name in ("cfg4", "cfg2"):      # the "+fp32" of SURVEY.md section 8 (the reference's iris examples run in fp32)
            r = bench_chains(ctx, name, k, wu, e2e=False, dtype="f32")
            if ctx.rank == 0:
                r["roofline"] = chain_roofline(ctx, r, WORKLOADS[name])
            extra[name + "_f32"] = strip(r)
        extra["cfg1"] = bench_cfg1(ctx)
        extra["cfg5"] = bench_datapar(ctx, k, wu, clocks=None, full=False)
    if ctx.rank == 0:
        if ctx.world == 1 and not args.no_cpu_baseline and w.get("kind") != "datapar":
            line["cpu_baseline"] = cpu_baseline_chains(wname)
        if extra:
            line["extra"] = extra
        print(json.dumps(line))
    ctx.close()


def cpu_baseline_chains(wname):
    """The oracle port on ONE host core, bounded sample; the unmodified reference itself cannot travel to the GPU box -- its
    timing on the build container's CPU is quoted from profiles/r02_reference_cpu_timing.json."""
    w = WORKLOADS[wname]
    cpp, it = cpu_sample_sizes(wname)
    cpp *= 16
    t = _cpu_worker((wname, cpp, it, 0))
    evals = cpp * it * (1 if w.get("kind") == "smmala" else w["num_steps"])
    cpu = {"value": evals / t, "unit": "evals/s", "cores": 1, "kind": "port",
           "sample": f"{cpp} chains x {it} iterations (L={w['num_steps']}) of the numpy oracle port, one process, {t:.1f} s"}
    ref_file = ROOT / "profiles" / "r02_reference_cpu_timing.json"
    if ref_file.exists():
        ref = json.loads(ref_file.read_text())
        cpu["note"] = ("the unmodified reference (torch autograd, /root/reference) cannot travel to the GPU box; timed in the build "
                       "container (%s, %d cores, one thread): cfg1 MALA.run(1100, 110) %.2f s = %.0f evals/s; HMC 2-3-2-1 XOR %.0f "
                       "evals/s; HMC 4-3-3 N=150 %.0f evals/s; 16-64-64-1 1M rows %.2f s per evaluation on %d threads "
                       "(tools/time_reference.py)" % (
                           ref["host"]["cpu"], ref["host"]["cores"], ref["cfg1_mala_221_1100_iters"]["seconds"],
                           ref["cfg1_mala_221_1100_iters"]["evals_per_s"], ref["cfg4_hmc_2321_xor"]["evals_per_s_counting_L"],
                           ref["cfg2_hmc_433_n150"]["evals_per_s_counting_L"],
                           ref["cfg5_upto_grad_16_64_64_1_n1M"]["seconds_per_eval_1M_rows"],
                           ref["cfg5_upto_grad_16_64_64_1_n1M"]["torch_threads"]))
    return cpu


def bench_datapar(ctx, steps, warmup, clocks=None, full=True):
    """BASELINE config 5: one replicated chain, rows sharded over the ranks (strong scaling); the 1 + P partial sums of every
    evaluation are exchanged by peer stores over NVLink inside the post kernel.  One step = `iters` HMC iterations (L
    evaluations each, every evaluation over ALL rows).  full=False: the compact sub-record of the default line."""
    torch, dist = ctx.torch, ctx.dist
    from torch.distributions import Normal

    from eeyore_b200.constants import loss_functions
    from eeyore_b200.models.mlp import MLP, Hyperparameters
    from eeyore_b200.samplers import DataShardedHMC, shard_rows

    args = ctx.args
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    w = WORKLOADS["cfg5"]
    n_total = args.rows or w["rows"]
    lo, hi = shard_rows(n_total, world, rank)
    gen = torch.Generator(device=dev).manual_seed(4)           # same stream on every rank; each keeps its slice
    teacher = torch.randn(16, device=dev, generator=gen)
    x = torch.empty(hi - lo, 16, device=dev)
    y = torch.empty(hi - lo, device=dev)
    chunk = 1 << 20
    for c0 in range(0, n_total, chunk):                         # generate the global data set chunk-wise, keep [lo, hi)
        c1 = min(n_total, c0 + chunk)
        xc = torch.randn(c1 - c0, 16, device=dev, generator=gen)
        yc = ((xc @ teacher + 0.5 * torch.randn(c1 - c0, device=dev, generator=gen)) > 0).float()
        a, b = max(lo, c0), min(hi, c1)
        if b > a:
            x[a - lo:b - lo], y[a - lo:b - lo] = xc[a - c0:b - c0], yc[a - c0:b - c0]
    hp = Hyperparameters(w["dims"], 3 * [True], 3 * [torch.sigmoid])
    model = MLP(loss=loss_functions[w["loss"]], hparams=hp, dtype=torch.float32, device=dev)
    P = model.num_params()
    model.prior = Normal(torch.zeros(P), S3 * torch.ones(P))
    iters, L = (args.iters if full else 0) or w["iters"], w["num_steps"]
    theta_host = (torch.randn(P, generator=torch.Generator().manual_seed(5)) * 0.1).pin_memory()

    sampler = DataShardedHMC(model, theta_host.to(dev), x, y, step=w["step"], num_steps=L, seed=7, exchange=args.exchange)
    for _ in range(warmup):
        sampler.run(num_epochs=iters, num_burnin_epochs=0)
    ctx.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ctx.barrier()
    if clocks:
        clocks.mark_begin()
    ev[0].record()
    for k in range(steps):
        sampler.run(num_epochs=iters, num_burnin_epochs=0)
        ev[k + 1].record()
    ctx.barrier()
    if clocks:
        clocks.mark_end()
        clocks.stop()
    t_res = ctx.max_over_ranks(ev[0].elapsed_time(ev[-1]) * 1e-3)
    per_step_ms = [round(ev[k].elapsed_time(ev[k + 1]), 2) for k in range(steps)]
    evals_step = iters * L
    value = evals_step * steps / t_res
    acc = sampler.acceptance_count() / max(1, sampler._iter)
    launches_per_iter = sampler.launches_per_iteration()
    launches_step = 1 if sampler.persistent else iters * launches_per_iter     # persistent: ONE cooperative launch per run()

    out_theta = torch.empty(iters, P).pin_memory()
    e2e_sampler = DataShardedHMC(model, theta_host, x, y, step=w["step"], num_steps=L, seed=11, exchange=args.exchange)

    def step_e2e():
        e2e_sampler.reset(theta_host)                        # chain state comes from the (pinned) host buffer
        samples, targets, accepted = e2e_sampler.run(num_epochs=iters, num_burnin_epochs=0)
        out_theta.copy_(samples, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    step_e2e()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_e2e()
    e1.record()
    ctx.barrier()
    t_e2e = ctx.max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    # ---- the dominant kernel alone (dp_eval_tc_kernel), CUDA events on the launching stream ---------------------------------
    import ctypes
    from eeyore_b200 import _native as nv
    lib = nv.lib()
    sums = torch.empty(P + 1, dtype=torch.float64, device=dev)
    ws = torch.empty(lib.eeyore_b200_dp_workspace_bytes() // 8, dtype=torch.float64, device=dev)
    th_dev = theta_host.to(dev)

    def time_kernel(fn, reps=20):
        for _ in range(3):
            nv.check(fn(nv.ptr(th_dev), nv.ptr(x), nv.ptr(y), hi - lo, nv.ptr(sums), nv.ptr(ws), nv.stream_ptr(dev)))
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(reps):
            nv.check(fn(nv.ptr(th_dev), nv.ptr(x), nv.ptr(y), hi - lo, nv.ptr(sums), nv.ptr(ws), nv.stream_ptr(dev)))
        k1.record()
        torch.cuda.synchronize()
        return k0.elapsed_time(k1) / reps

    amax = torch.zeros(1, dtype=torch.float32, device=dev)
    nv.check(lib.eeyore_b200_dp_absmax(nv.ptr(x), x.numel(), nv.ptr(amax), nv.stream_ptr(dev)))

    def tc_kernel(th, xx, yy, nn, out, wsp, st):          # max |x| of the shard computed once, as DataShardedHMC does
        return lib.eeyore_b200_dp_loglik_grad_x(th, xx, yy, nn, nv.ptr(amax), out, wsp, st)

    ctx.barrier()
    kernel_ms = ctx.max_over_ranks(time_kernel(tc_kernel))
    ffma_ms = ctx.max_over_ranks(time_kernel(lib.eeyore_b200_dp_loglik_grad_ffma, reps=5)) if full else None
    line = None
    if rank == 0:
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
            tensor_peak, peak_src = float(peaks["bf16_tflops"]), "MEASURED_PEAKS.json bf16_tflops (burst; the kernel is timed alone)"
        except Exception:
            tensor_peak, peak_src = 1590.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
        achieved = 29056.0 * (hi - lo) / (kernel_ms * 1e-3) / 1e12
        # fp16 MMA work actually issued per 128-row tile: three piece products per GEMM as two MMAs (N = 128 and 64, plus the
        # 8 ones columns of the weight-gradient GEMMs); the M = 64 weight-gradient MMAs run at the cost of M = 128
        tiles = (hi - lo + 127) // 128
        mma_flops_tile = 2 * 128 * 16 * (128 + 64) + 2 * (2 * 128 * 64 * (128 + 64)) \
            + 2 * 128 * 128 * (136 + 72) + 2 * 128 * 128 * (40 + 24)
        exchange = ("none (1 GPU)" if world == 1 else
                    "NCCL all-reduce of 1+P fp64 sums per evaluation" if sampler.exchange == "nccl" else
                    "1+P fp64 sums stored into every peer's inbox over NVLink (CUDA IPC), sequence-numbered flags, totals added in "
                    "rank order; no NCCL on the data path")
        roofline = {"bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                    "traffic": _cfg5_traffic(hi - lo), "peak_source": peak_src,
                    "kernel": "dp_eval_tc_kernel (tcgen05 fp16 MMA on exact two-piece splits, fp32 accumulate in TMEM, two tiles in flight)",
                    "avg_launch_ms": kernel_ms, "algorithmic_flops_per_row": 29056,
                    "f16_mma_tflops_issued": tiles * mma_flops_tile / (kernel_ms * 1e-3) / 1e12,
                    "share_of_step": kernel_ms * evals_step / (1e3 * t_res / steps),
                    "hbm_view": {"algorithmic_bytes_per_launch": (hi - lo) * 68, "achieved_gbs": (hi - lo) * 68 / (kernel_ms * 1e-3) / 1e9}}
        config = {"workload": w["name"], "rows_total": n_total, "rows_per_gpu": hi - lo, "hmc_iterations_per_step": iters,
                  "num_steps": L, "evals_counted_per_iteration": L, "exchange": exchange, "trajectory": sampler.mode,
                  "launches_per_hmc_iteration": launches_per_iter, "acceptance_rate": acc,
                  "l2": "x shard (%.0f MB) exceeds the 126 MB L2" % ((hi - lo) * 68 / 1e6),
                  "data_resident": "x, y shards stay in HBM across steps; e2e copies the chain state in and the samples out"}
        e2e = {"value": evals_step * steps / t_e2e, "unit": "evals/s", "h2d_bytes_per_step": P * 4, "d2h_bytes_per_step": iters * P * 4,
               "ms_per_step": 1e3 * t_e2e / steps}
        if full:
            peak32 = ctypes.c_double()
            nv.check(lib.eeyore_b200_fma_peak(nv.F32, 4000, ctypes.byref(peak32)))
            roofline["note"] = ("fp32 parity costs three fp16 piece products per GEMM (and M = 64 padding): the tensor pipe executes "
                                "%.1fx the algorithmic FLOPs" % (mma_flops_tile / (29056.0 * 128)))
            roofline["fp32_fma_view"] = {"peak": peak32.value, "frac": achieved / peak32.value,
                                         "peak_source": "measured live by eeyore_b200_fma_peak (this device)",
                                         "ffma_kernel_ms": ffma_ms, "speedup_over_ffma_kernel": ffma_ms / kernel_ms}
            config["per_step_ms"] = per_step_ms
            cpu = None
            if world == 1 and not args.no_cpu_baseline:
                procs = host_cores()
                pool = CpuPool(procs)
                cpu_datapar_throughput(pool, procs, 1 << 16, n_total)
                v, wall = cpu_datapar_throughput(pool, procs, 1 << 20, n_total)
                pool.close()
                cpu = {"value": v, "unit": "evals/s", "cores": procs, "kind": "port",
                       "sample": f"one evaluation over {1 << 20} of the {n_total} rows split over {procs} processes of the numpy oracle "
                                 f"port, {wall:.1f} s, scaled linearly to {n_total} rows"}
            line = {"metric": "log_target_grad_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world, "steps": steps,
                    "warmup": warmup, "ms_per_step": 1e3 * t_res / steps, "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "e2e": e2e,
                    "gpu_launches": steps * launches_step, "clocks": clocks.summary() if clocks else None,
                    "roofline": roofline, "cpu_baseline": cpu}
        else:
            line = {"value": value, "unit": "evals/s (each over all %d rows)" % n_total, "ms_per_step": 1e3 * t_res / steps,
                    "ms_per_evaluation": 1e3 * t_res / steps / evals_step, "scaling": "strong", "dtype": "f32", "steps": steps,
                    "config": config, "e2e": e2e, "kernel_ms": kernel_ms, "gpu_launches": steps * launches_step,
                    "roofline": roofline}
    sampler.check_status()
    sampler.close()
    e2e_sampler.close()
    del x, y
    torch.cuda.empty_cache()
    return line


def _cfg5_traffic(rows_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of dp_eval_tc_kernel from the committed ncu --set full capture; only when
    this launch processes the row count the capture was taken at (one GPU, 8,388,608 rows), else null."""
    tfile = ROOT / "profiles" / "r01_traffic_cfg5.json"
    if not tfile.exists():
        return None
    rec = json.loads(tfile.read_text())
    return rec["dram_bytes_total"] if rec.get("rows_per_launch") == rows_per_launch else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=list(WORKLOADS))
    ap.add_argument("--dtype", default=None, choices=["f32", "f64"], help="chain workloads: override the arithmetic type")
    ap.add_argument("--chains", type=int, default=0, help="override chains per GPU")
    ap.add_argument("--iters", type=int, default=0, help="override HMC iterations per step")
    ap.add_argument("--rows", type=int, default=0, help="cfg5: override the total number of data rows")
    ap.add_argument("--lanes", type=int, default=0, help="threads cooperating on one chain (0 = library heuristic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="default cfg4 line only, without the sub-records of the other BASELINE configurations")
    ap.add_argument("--diag-samples", type=int, default=1000, help="saved iterations per chain of the diagnostics stage")
    ap.add_argument("--diag-chunk", type=int, default=75776,
                    help="chains per chunk of the diagnostics stage (default: two full waves of the fp64 chain kernel, 2 x 148 SMs x 2 CTAs "
                         "x 128 chains; 65,536 chains are 1.73 waves and take as long)")
    ap.add_argument("--e2e-batches", type=int, default=4,
                    help="chain batches (streams) of the end-to-end arm: copies of one batch overlap the kernel of another")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="cfg5: exchange step of the data-sharded path (auto = peer stores over NVLink when there are several ranks)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
