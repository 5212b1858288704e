#!/usr/bin/env python
"""Benchmark of the sampler hot path (BASELINE.json metric: log-target+gradient evaluations per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg2] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

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
def run_ours(args):
    import torch
    import torch.distributed as dist
    from torch.distributions import Normal
    from torch.utils.data import DataLoader

    from eeyore_b200 import _native as nv
    from eeyore_b200.constants import loss_functions
    from eeyore_b200.datasets import XYDataset
    from eeyore_b200.models.mlp import MLP, Hyperparameters
    from eeyore_b200.samplers import HMC

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = WORKLOADS[args.workload]
    dt = torch.float64 if w["dtype"] == "f64" else torch.float32
    x, y = synthetic_data(w)
    ds = XYDataset(torch.from_numpy(x).to(dt), torch.from_numpy(y).to(dt))
    nl = len(w["dims"]) - 1
    binary = w["loss"] == "binary_classification"
    hp = Hyperparameters(w["dims"], nl * [True], (nl - 1) * [torch.sigmoid] + [torch.sigmoid if binary else None])
    model = MLP(loss=loss_functions[w["loss"]], hparams=hp, dtype=dt, device=dev)
    P = model.num_params()
    model.prior = Normal(torch.zeros(P, dtype=dt), S3 * torch.ones(P, dtype=dt))
    C = args.chains or w["chains"]
    iters, L, thin = args.iters or w["iters"], w["num_steps"], w["thin"]
    gen = torch.Generator().manual_seed(1000 + rank)
    theta_host = (torch.randn(C, P, generator=gen, dtype=dt) * {"xor": 1.0, "noisy_xor": 0.5}.get(w["data"], 0.3)).pin_memory()
    loader = DataLoader(ds, batch_size=len(ds))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- device-resident arm: state stays in HBM, one fused launch per step --------------------------------------
    kind = w.get("kind", "hmc")

    def make_sampler(theta0, seed):
        if kind == "smmala":
            from eeyore_b200.samplers import SMMALA
            return SMMALA(model, theta0=theta0, dataloader=loader, step=w["step"], seed=seed, thin=thin)
        return HMC(model, theta0=theta0, dataloader=loader, step=w["step"], num_steps=L, seed=seed, thin=thin,
                   lanes_per_chain=args.lanes)

    sampler = make_sampler(theta_host.to(dev), 12345)
    sampler.chain_offset = rank * C          # global chain ids => results independent of the sharding

    def step_resident():
        sampler._device_blocks = []          # saved states of the previous step are dropped (buffer is recycled)
        sampler.counter.reset()
        sampler.run(num_epochs=iters, num_burnin_epochs=0)

    clocks = ClockSampler(local, enabled=(rank == 0)).start()
    for _ in range(args.warmup):
        step_resident()
    barrier()
    # Workloads whose per-launch working set fits the 126 MB L2 (configs 2 and 3) get the L2 flushed between timed steps (a
    # 160 MB buffer is overwritten; the flush sits outside the per-step event pairs); config 4 streams 789 MB per launch.
    esz0 = theta_host.element_size()
    ws_bytes = C * (2 * (2 * P + 1) * esz0 + ((iters + thin - 1) // thin) * (P * esz0 + esz0 + 1) + 4)
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev) if ws_bytes <= 126e6 else None
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    clocks.mark_begin()
    for k in range(args.steps):
        if flush is not None:
            flush.fill_(k & 0xFF)
        ev0[k].record()
        step_resident()
        ev1[k].record()
    barrier()
    clocks.mark_end()
    clocks.stop()
    per_launch_ms = [ev0[k].elapsed_time(ev1[k]) for k in range(args.steps)]
    t_local = (sum(per_launch_ms) if flush is not None else ev0[0].elapsed_time(ev1[-1])) * 1e-3
    t_res = max_over_ranks(t_local)
    evals_step = C * iters * (1 if kind == "smmala" else L)
    value = world * evals_step * args.steps / t_res
    acc_rate = sampler.get_chain().acceptance().mean().item()

    # ---- end-to-end arm: host buffers in, host results out, through the public API --------------------------------
    out_theta = torch.empty(C, P, dtype=dt).pin_memory()
    out_lt = torch.empty(C, dtype=dt).pin_memory()
    out_acc = torch.empty(C, dtype=torch.int32).pin_memory()

    # The chains are independent, so the step is cut into `nb` chain batches, each on its own stream: batch b + 1's
    # host->device copy and batch b - 1's device->host copy run under batch b's kernel.  Philox is keyed by the global chain
    # id, so the results do not depend on the batching.
    nb = max(1, min(args.e2e_batches, C // 1024 if C >= 1024 else 1))
    streams = [torch.cuda.Stream(device=dev) for _ in range(nb)]
    bounds = [(b * C // nb, (b + 1) * C // nb) for b in range(nb)]

    def step_e2e():
        cur = torch.cuda.current_stream()
        keep = []
        for (lo, hi), st in zip(bounds, streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                s = make_sampler(theta_host[lo:hi], 999)
                s.chain_offset = rank * C + lo
                s.run(num_epochs=iters, num_burnin_epochs=0)
                out_theta[lo:hi].copy_(s.current["sample"], non_blocking=True)
                out_lt[lo:hi].copy_(s.current["target_val"], non_blocking=True)
                out_acc[lo:hi].copy_(s.acceptance_counts(), non_blocking=True)
            keep.append(s)                       # buffers stay alive until their stream has drained
        for st in streams:
            cur.wait_stream(st)
        cur.synchronize()
        return out_lt[0].item()

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    t_e2e = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    e2e_value = world * evals_step * args.steps / t_e2e
    h2d = (theta_host.numel() + x.size + y.size) * theta_host.element_size()
    d2h = sum(t.numel() * t.element_size() for t in (out_theta, out_lt, out_acc))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (sampler_kernel<..., KIND_HMC>) -------------------------------------------
    import ctypes
    peak = ctypes.c_double()
    nv.check(nv.lib().eeyore_b200_fma_peak(nv.DTYPE_IDS[dt], 2000, ctypes.byref(peak)))
    avg_launch_s = float(np.mean(per_launch_ms)) * 1e-3
    flops_launch = w["flops_per_eval"] * evals_step
    achieved = flops_launch / avg_launch_s / 1e12
    esz = theta_host.element_size()
    n_saved = (iters + thin - 1) // thin
    hbm_bytes = C * (2 * (2 * P + 1) * esz + n_saved * (P * esz + esz + 1) + 4)
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    hbm_peak = json.loads(peaks_file.read_text())["hbm_gbs"] if peaks_file.exists() else 6650.0
    traffic = None
    tfile = ROOT / "profiles" / "r01_traffic_cfg4.json"
    if args.workload == "cfg4" and not args.chains and not args.iters and tfile.exists():
        traffic = json.loads(tfile.read_text())["dram_bytes_total"]   # dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full
    roofline = {"bound": "fp64_fma" if w["dtype"] == "f64" else "fp32_fma", "achieved": achieved, "peak": peak.value,
                "unit": "TFLOP/s", "frac": achieved / peak.value, "traffic": traffic,
                "peak_source": "measured live by eeyore_b200_fma_peak (dependent-free FMA chains, this device)",
                "algorithmic_flops_per_eval": w["flops_per_eval"], "avg_launch_ms": avg_launch_s * 1e3,
                "hbm_view": {"algorithmic_bytes_per_launch": hbm_bytes, "achieved_gbs": hbm_bytes / avg_launch_s / 1e9,
                             "peak_gbs": hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json" if peaks_file.exists() else "fallback"}}

    # ---- CPU baseline: the oracle port on one core, bounded sample -------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpp, it = cpu_sample_sizes(args.workload)
        cpp *= 16
        t = _cpu_worker((args.workload, cpp, it, 0))
        cpu = {"value": cpp * it * (1 if kind == "smmala" else L) / t, "unit": "evals/s", "cores": 1, "kind": "port",
               "sample": f"{cpp} chains x {it} HMC iterations (L={L}) of the numpy oracle port, one process, {t:.1f} s"}

    print(json.dumps({
        "metric": "log_target_grad_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
        "config": {"workload": w["name"], "chains_per_gpu": C, "hmc_iterations_per_step": iters, "num_steps": L,
                   "step_size": w["step"], "thin": thin, "evals_counted_per_iteration": L,
                   "evals_reference_executes_per_iteration": L + 1, "rng": "philox4x32-10 on device",
                   "acceptance_rate": acc_rate,
                   "l2": ("chain state + saved samples per launch (%.0f MB) exceed the 126 MB L2" % (hbm_bytes / 1e6))
                         if hbm_bytes > 126e6 else
                         ("chain state + saved samples per launch are %.0f MB (below the 126 MB L2): 160 MB of device memory "
                          "are overwritten between timed steps to flush it" % (hbm_bytes / 1e6))},
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * t_e2e / args.steps, "chain_batches": nb,
                "note": "public sampler API on %d chain batches / streams: each batch copies its theta in from pinned host "
                        "memory, runs, and copies states / targets / accept counts out; copies overlap other batches' kernels" % nb},
        "gpu_launches": args.steps,
        "gpu_launches_e2e": 2 * nb * args.steps,
        "clocks": clocks.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }))
    if world > 1:
        dist.destroy_process_group()


def run_datapar(args):
    """BASELINE config 5: one replicated chain, rows sharded over the ranks (strong scaling), NCCL all-reduce of the
    1 + P partial sums after every local evaluation.  One step = `iters` HMC iterations (L evaluations each, every
    evaluation over ALL rows)."""
    import torch
    import torch.distributed as dist
    from torch.distributions import Normal

    from eeyore_b200.constants import loss_functions
    from eeyore_b200.models.mlp import MLP, Hyperparameters
    from eeyore_b200.samplers import DataShardedHMC, shard_rows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOADS[args.workload]
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
    iters, L = args.iters or w["iters"], w["num_steps"]
    theta_host = (torch.randn(P, generator=torch.Generator().manual_seed(5)) * 0.1).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    sampler = DataShardedHMC(model, theta_host.to(dev), x, y, step=w["step"], num_steps=L, seed=7, exchange=args.exchange)
    clocks = ClockSampler(local, enabled=(rank == 0 and not os.environ.get("EEYORE_BENCH_NO_CLOCKS"))).start()
    for _ in range(args.warmup):
        sampler.run(num_epochs=iters, num_burnin_epochs=0)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    clocks.mark_begin()
    ev[0].record()
    for k in range(args.steps):
        sampler.run(num_epochs=iters, num_burnin_epochs=0)
        ev[k + 1].record()
    barrier()
    clocks.mark_end()
    clocks.stop()
    t_res = max_over_ranks(ev[0].elapsed_time(ev[-1]) * 1e-3)
    per_step_ms = [round(ev[k].elapsed_time(ev[k + 1]), 2) for k in range(args.steps)]
    evals_step = iters * L
    value = evals_step * args.steps / t_res
    acc = sampler.acceptance_count() / max(1, sampler._iter)

    out_theta = torch.empty(iters, P).pin_memory()

    e2e_sampler = DataShardedHMC(model, theta_host, x, y, step=w["step"], num_steps=L, seed=11, exchange=args.exchange)

    def step_e2e():
        e2e_sampler.reset(theta_host)                        # chain state comes from the (pinned) host buffer
        samples, targets, accepted = e2e_sampler.run(num_epochs=iters, num_burnin_epochs=0)
        out_theta.copy_(samples, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    t_e2e = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    # ---- the dominant kernel alone (dp_eval_tc_kernel + its 21-CTA reduction), CUDA events on the launching stream -------
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

    barrier()
    kernel_ms = max_over_ranks(time_kernel(tc_kernel))
    ffma_ms = max_over_ranks(time_kernel(lib.eeyore_b200_dp_loglik_grad_ffma, reps=5))
    if rank == 0:
        peak32 = ctypes.c_double()
        nv.check(lib.eeyore_b200_fma_peak(nv.F32, 4000, ctypes.byref(peak32)))
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
            tensor_peak, peak_src = float(peaks["bf16_tflops"]), "MEASURED_PEAKS.json bf16_tflops (burst; the kernel is timed alone)"
        except Exception:
            tensor_peak, peak_src = 1590.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
        flops_launch = 29056.0 * (hi - lo)
        achieved = flops_launch / (kernel_ms * 1e-3) / 1e12
        # fp16 MMA work actually issued per 128-row tile: three piece products per GEMM as two MMAs (N = 128 and 64, plus the
        # 8 ones columns of the weight-gradient GEMMs); the M = 64 weight-gradient MMAs run at the cost of M = 128
        tiles = (hi - lo + 127) // 128
        mma_flops_tile = 2 * 128 * 16 * (128 + 64) + 2 * (2 * 128 * 64 * (128 + 64)) \
            + 2 * 128 * 128 * (136 + 72) + 2 * 128 * 128 * (40 + 24)
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
        print(json.dumps({
            "metric": "log_target_grad_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "rows_total": n_total, "rows_per_gpu": hi - lo,
                       "hmc_iterations_per_step": iters, "num_steps": L, "evals_counted_per_iteration": L,
                       "exchange": ("none (1 GPU)" if world == 1 else
                                    "NCCL all-reduce of 1+P fp64 sums per evaluation" if sampler.exchange == "nccl" else
                                    "1+P fp64 sums stored into every peer's inbox over NVLink (CUDA IPC) inside the fused post "
                                    "kernel, sequence-numbered flags, totals added in rank order; no NCCL on the data path"),
                       "acceptance_rate": acc, "per_step_ms": per_step_ms,
                       "l2": "x shard (%.0f MB) exceeds the 126 MB L2" % ((hi - lo) * 68 / 1e6),
                       "data_resident": "x, y shards stay in HBM across steps; e2e copies the chain state in and the samples out"},
            "e2e": {"value": evals_step * args.steps / t_e2e, "unit": "evals/s", "h2d_bytes_per_step": P * 4,
                    "d2h_bytes_per_step": iters * P * 4, "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": args.steps * iters * (2 + (4 if sampler.exchange == "nccl" else 2) * L),
            "clocks": clocks.summary(),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                         "frac": achieved / tensor_peak, "traffic": _cfg5_traffic(hi - lo), "peak_source": peak_src,
                         "kernel": "dp_eval_tc_kernel (tcgen05 fp16 MMA on exact two-piece splits, fp32 accumulate in TMEM, two tiles in flight)",
                         "avg_launch_ms": kernel_ms, "algorithmic_flops_per_row": 29056,
                         "f16_mma_tflops_issued": tiles * mma_flops_tile / (kernel_ms * 1e-3) / 1e12,
                         "note": "fp32 parity costs three fp16 piece products per GEMM (and M = 64 padding): the tensor pipe "
                                 "executes %.1fx the algorithmic FLOPs" % (mma_flops_tile / (29056.0 * 128)),
                         "fp32_fma_view": {"peak": peak32.value, "frac": achieved / peak32.value,
                                           "peak_source": "measured live by eeyore_b200_fma_peak (this device)",
                                           "ffma_kernel_ms": ffma_ms, "speedup_over_ffma_kernel": ffma_ms / kernel_ms},
                         "share_of_step": kernel_ms * evals_step / (1e3 * t_res / args.steps),
                         "hbm_view": {"algorithmic_bytes_per_launch": (hi - lo) * 68,
                                      "achieved_gbs": (hi - lo) * 68 / (kernel_ms * 1e-3) / 1e9}},
            "cpu_baseline": cpu,
        }))
    sampler.check_status()
    sampler.close()
    e2e_sampler.close()
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--chains", type=int, default=0, help="override chains per GPU")
    ap.add_argument("--iters", type=int, default=0, help="override HMC iterations per step")
    ap.add_argument("--rows", type=int, default=0, help="cfg5: override the total number of data rows")
    ap.add_argument("--lanes", type=int, default=0, help="threads cooperating on one chain (0 = library heuristic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-batches", type=int, default=8,
                    help="chain batches (streams) of the end-to-end arm: copies of one batch overlap the kernel of another")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="cfg5: exchange step of the data-sharded path (auto = peer stores over NVLink when there are several ranks)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif WORKLOADS[args.workload].get("kind") == "datapar":
        run_datapar(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
