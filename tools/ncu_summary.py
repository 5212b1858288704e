"""Condenses the CSV pages of an ncu report (exported on the GPU box with `ncu -i rep --page raw|source --csv`) into the few
numbers the design notes quote.   python tools/ncu_summary.py raw.csv [src.csv.gz] [--top 25]"""
import csv
import gzip
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("== kernel:", d.get("Kernel Name", "")[:100])
        for k in KEYS:
            if k in d:
                print(f"  {k:85s} {d[k]}")
        stalls = sorted(((float(d[k] or 0), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                         for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio")), reverse=True)
        print("  stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in stalls[:8]))


def source(path, top):
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.reader(op(path, "rt")))
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    samp = next((c for c in ("# Samples", "Samples", "Warp Stall Sampling (All Samples)") if c in col), None)
    src_c = next((c for c in ("Source", "Source Line") if c in col), None)
    print("columns:", [h for h in hdr][:40])
    if samp is None:
        return
    by_line = defaultdict(float)
    total = 0.0
    for r in rows[1:]:
        try:
            v = float(r[col[samp]] or 0)
        except ValueError:
            continue
        total += v
        by_line[r[col[src_c]] if src_c else r[0]] += v
    for line, v in sorted(by_line.items(), key=lambda kv: -kv[1])[:top]:
        print(f"  {100 * v / max(total, 1):5.1f}%  {line[:150]}")


if __name__ == "__main__":
    top = 25
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if "--top" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1])
        args = [a for a in args if a != str(top)]
    raw(args[0])
    if len(args) > 1:
        source(args[1], top)
