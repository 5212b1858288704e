"""Event timeline of bench.py's end-to-end step (cfg4): where the time between `value` and `e2e` goes.
   python tools/e2e_timeline.py [n_batches ...]      (GPU box)"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("nbs", nargs="*", type=int, default=[8])
ap.add_argument("--no-copy-out", action="store_true")
ap.add_argument("--zero-copy", action="store_true", help="saved states written by the kernel straight into pinned host memory")
ap.add_argument("--waves", type=str, default=None, help="comma list: waves of the first batches / last batches, 'r' = the rest, e.g. 1,r,2")
ap.add_argument("--pipeline", type=int, nargs="*", default=None, help="waves per batch: h2d / compute / d2h streams, batches in order")
a = ap.parse_args()
args = argparse.Namespace(gpus=1, chains=0, iters=0, lanes=0, e2e_batches=8)
ctx = bench.Ctx(args)
w = dict(bench.WORKLOADS["cfg4"])
model, loader, dt, x, y = bench.build_model(ctx, w)
P, C, iters, thin = model.num_params(), w["chains"], w["iters"], w["thin"]
n_saved = (iters + thin - 1) // thin
theta_host = (torch.randn(C, P, dtype=dt)).pin_memory()
pin = lambda *shape, dtype=dt: torch.empty(*shape, dtype=dtype).pin_memory()
for nb in (a.nbs if a.pipeline is None else []):
    bounds = [(b * C // nb, (b + 1) * C // nb) for b in range(nb)]
    if a.waves:
        WAVE = 148 * 2 * 128
        spec = a.waves.split(",")
        fixed = sum(int(float(t) * WAVE) for t in spec if t != "r")
        sizes = [C - fixed if t == "r" else int(float(t) * WAVE) for t in spec]
        edges = [0]
        for z in sizes:
            edges.append(edges[-1] + z)
        bounds = list(zip(edges[:-1], edges[1:]))
        nb = len(bounds)
    streams = [torch.cuda.Stream() for _ in range(nb)]
    samplers = []
    for lo, hi in bounds:
        s = bench.make_sampler(w, model, loader, theta_host[lo:hi], 999, thin, 1)
        s.chain_offset = lo
        s.host_output = a.zero_copy
        samplers.append(s)
    out_samples = [pin(n_saved, P, hi - lo) for lo, hi in bounds]
    out_targets = [pin(n_saved, hi - lo) for lo, hi in bounds]
    out_accepted = [pin(n_saved, hi - lo, dtype=torch.uint8) for lo, hi in bounds]
    out_theta, out_lt, out_acc = pin(C, P), pin(C), pin(C, dtype=torch.int32)

    def step(record):
        cur = torch.cuda.current_stream()
        ev = []
        e_begin = torch.cuda.Event(enable_timing=True)
        e_begin.record()
        for b, ((lo, hi), st, s) in enumerate(zip(bounds, streams, samplers)):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                es = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                es[0].record()
                s.reset(theta_host[lo:hi])
                es[1].record()
                s.run(num_epochs=iters, num_burnin_epochs=0)
                es[2].record()
                if not a.no_copy_out:
                    blk = s._device_blocks[-1]
                    if not a.zero_copy:
                        out_samples[b].copy_(blk["sample"], non_blocking=True)
                        out_targets[b].copy_(blk["target_val"], non_blocking=True)
                        out_accepted[b].copy_(blk["accepted"], non_blocking=True)
                    out_theta[lo:hi].copy_(s.current["sample"], non_blocking=True)
                    out_lt[lo:hi].copy_(s.current["target_val"], non_blocking=True)
                    out_acc[lo:hi].copy_(s.acceptance_counts(), non_blocking=True)
                es[3].record()
                ev.append(es)
        for st in streams:
            cur.wait_stream(st)
        e_end = torch.cuda.Event(enable_timing=True)
        e_end.record()
        cur.synchronize()
        if record:
            rows = [[round(e_begin.elapsed_time(e), 2) for e in es] for es in ev]
            return {"nb": nb, "total_ms": round(e_begin.elapsed_time(e_end), 2), "per_batch [start, reset done, run done, copies done]": rows}

    for _ in range(3):
        step(False)
    import time
    t0 = time.perf_counter()
    r = step(True)
    r["wall_ms"] = round(1e3 * (time.perf_counter() - t0), 2)
    print(json.dumps(r))
    del samplers
    torch.cuda.empty_cache()


if a.pipeline is not None:
    WAVE = 148 * 2 * 128
    for wpb in a.pipeline:
        bounds, lo = [], 0
        while lo < C:
            hi = min(C, lo + wpb * WAVE)
            if C - hi < WAVE // 2:      # fold a small remainder into the last batch
                hi = C
            bounds.append((lo, hi))
            lo = hi
        if wpb > 1 and bounds[-1][1] - bounds[-1][0] > WAVE:      # short tail: split the last batch at a wave boundary
            lo, hi = bounds.pop()
            bounds += [(lo, lo + WAVE * ((hi - lo) // WAVE)), (lo + WAVE * ((hi - lo) // WAVE), hi)] if (hi - lo) % WAVE else [(lo, hi)]
        nb = len(bounds)
        s_in, s_out, s_c = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        samplers, stage = [], []
        for lo, hi in bounds:
            s = bench.make_sampler(w, model, loader, theta_host[lo:hi], 999, thin, 1)
            s.chain_offset = lo
            samplers.append(s)
            stage.append(torch.empty(hi - lo, P, dtype=dt, device="cuda"))
        out_samples = [pin(n_saved, P, hi - lo) for lo, hi in bounds]
        out_targets = [pin(n_saved, hi - lo) for lo, hi in bounds]
        out_accepted = [pin(n_saved, hi - lo, dtype=torch.uint8) for lo, hi in bounds]
        out_theta, out_lt, out_acc = pin(C, P), pin(C), pin(C, dtype=torch.int32)

        def step2(record):
            cur = torch.cuda.current_stream()
            e_begin = torch.cuda.Event(enable_timing=True)
            e_begin.record()
            for st in (s_in, s_c, s_out):
                st.wait_stream(cur)
            ev = []
            for b, ((lo, hi), s) in enumerate(zip(bounds, samplers)):
                es = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                with torch.cuda.stream(s_in):
                    stage[b].copy_(theta_host[lo:hi], non_blocking=True)
                    es[0].record()
                with torch.cuda.stream(s_c):
                    s_c.wait_event(es[0])
                    s.reset(stage[b])
                    s.run(num_epochs=iters, num_burnin_epochs=0)
                    es[1].record()
                with torch.cuda.stream(s_out):
                    s_out.wait_event(es[1])
                    blk = s._device_blocks[-1]
                    out_samples[b].copy_(blk["sample"], non_blocking=True)
                    out_targets[b].copy_(blk["target_val"], non_blocking=True)
                    out_accepted[b].copy_(blk["accepted"], non_blocking=True)
                    out_theta[lo:hi].copy_(s.current["sample"], non_blocking=True)
                    out_lt[lo:hi].copy_(s.current["target_val"], non_blocking=True)
                    out_acc[lo:hi].copy_(s.acceptance_counts(), non_blocking=True)
                    es[2].record()
                ev.append(es)
            for st in (s_in, s_c, s_out):
                cur.wait_stream(st)
            e_end = torch.cuda.Event(enable_timing=True)
            e_end.record()
            cur.synchronize()
            if record:
                return {"pipeline_waves_per_batch": wpb, "nb": nb, "sizes": [hi - lo for lo, hi in bounds], "total_ms": round(e_begin.elapsed_time(e_end), 2),
                        "per_batch [h2d done, run done, d2h done]": [[round(e_begin.elapsed_time(e), 2) for e in es] for es in ev]}

        for _ in range(3):
            step2(False)
        import time
        t0 = time.perf_counter()
        r = step2(True)
        r["wall_ms"] = round(1e3 * (time.perf_counter() - t0), 2)
        print(json.dumps(r))
        del samplers
        torch.cuda.empty_cache()
