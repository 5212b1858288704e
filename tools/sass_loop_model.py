#!/usr/bin/env python3
"""Static model of the fp64 sampler kernel's row loop, used while tuning it without a GPU (round 1).

   cuobjdump -sass kernel.cubin > k.sass;  [MINRCP=6] [SKIP=first-last] python tools/sass_loop_model.py k.sass [warps] [dfma_latency]

Parses the SASS, picks the smallest loop that holds at least MINRCP MUFU.RCP64H (one per sigmoid), follows unconditional
forward branches (SKIP drops an instruction index range, e.g. the out-of-line general path), and simulates `warps` warps
issuing in order on one scheduler: one issue per cycle, register scoreboard, an FP64 pipe that an instruction occupies for
2.4 cycles, latencies calibrated against ncu warp-stall samples on a B200 (dependent DFMA 11.8 cycles, LDS 44, MUFU 22).
It predicted the direction of every change that was then measured (e.g. 1908 -> 1324 cycles per row for the instruction
diet of the sigmoid / log, measured 1897 -> 1497) but is optimistic about what the compiler's interleaving buys.
Development aid only: nothing in the product or the tests uses it."""
import re, sys
LAT_D = 11.8
import os
MINRCP = int(os.environ.get("MINRCP", "6"))
W = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lines = []
for l in open(sys.argv[1]):
    m = re.search(r'^\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if m: lines.append((int(m.group(1), 16), m.group(2).strip()))
addr2i = {a: i for i, (a, _) in enumerate(lines)}
loops = []
for i, (a, t) in enumerate(lines):
    m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(?:`\()?(0x[0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr2i: loops.append((addr2i[tgt], i))
best = None
for s, e in loops:
    n = sum(1 for k in range(s, e + 1) if 'MUFU.RCP64H' in lines[k][1])
    if n >= MINRCP and (best is None or e - s < best[1] - best[0]): best = (s, e)
s, e = best
SKIP = os.environ.get('SKIP')
skip = tuple(int(x) for x in SKIP.split('-')) if SKIP else None
body = []
k = s
while k <= e:
    t = lines[k][1]
    m = re.match(r'BRA\s+(0x[0-9a-f]+)', t)
    if m and int(m.group(1),16) in addr2i and k < addr2i[int(m.group(1),16)] <= e:
        k = addr2i[int(m.group(1),16)]; continue
    if skip and skip[0] <= k <= skip[1]:
        k += 1; continue
    body.append(t); k += 1
# drop the soft-label branch region: instructions after an unconditional BRA until its target? keep simple: stop region skipping
def regs(tok, wide):
    out = []
    for m in re.finditer(r'(?<![A-Za-z0-9_])(U?R)(\d+)(?:\.64)?', tok):
        n = int(m.group(2)); p = m.group(1)
        out.append((p, n))
        if wide: out.append((p, n + 1))
    return out
def parse(t):
    t = re.sub(r'^@!?U?P\d+\s+', '', t)
    op, _, rest = t.partition(' ')
    base = op.split('.')[0]
    ops = [x.strip() for x in rest.split(',')] if rest else []
    isD = base in ('DFMA', 'DADD', 'DMUL', 'DSETP')
    wide_dst = isD and base != 'DSETP' or '.64' in op or base in ('I2F',) and 'F64' in op
    w128 = '.128' in op
    dst, src = [], []
    if base in ('STS', 'STG', 'STL', 'BRA', 'BSYNC', 'BSSY', 'EXIT', 'BAR', 'NOP', 'WARPSYNC'):
        for o in ops: src += regs(o, '.64' in op)
    else:
        if ops:
            d = regs(ops[0], wide_dst)
            if w128 and d: d = [(d[0][0], d[0][1] + k) for k in range(4)]
            if base == 'DSETP' or base == 'ISETP' or base == 'FSETP': d = []  # predicates ignored (approx)
            dst = d
        for o in ops[1:]:
            src += regs(o, isD)
        if base == 'MUFU' and 'RCP64H' in op: pass
    return base, op, dst, src
P = [parse(t) for t in body]
def lat(base, op):
    if base in ('DFMA', 'DADD', 'DMUL', 'DSETP'): return LAT_D
    if base == 'MUFU': return 22
    if base in ('LDS','LDL'): return 44
    if base in ('LDC', 'LDCU'): return 30
    if base == 'I2F': return 14
    return 5
def occ(base):
    return 2.4 if base in ('DFMA', 'DADD', 'DMUL', 'DSETP') else (2.0 if base=='MUFU' else 1.0)
nD = sum(1 for b, *_ in P if b in ('DFMA', 'DADD', 'DMUL', 'DSETP'))
print(f'loop lines {s}..{e}: {len(P)} instrs, {nD} FP64')
# simulate
ITER = 6
ready = [dict() for _ in range(W)]
pc = [0] * W; it = [0] * W; tnext = [w * 7.0 for w in range(W)]
pipe_free = 0.0; now = 0.0; issue_free = 0.0
done = [None] * W; start = [None]*W
import heapq
while any(it[w] < ITER for w in range(W)):
    # pick the warp that can issue earliest
    cand = []
    for w in range(W):
        if it[w] >= ITER: continue
        base, op, dst, src = P[pc[w]]
        t = max(tnext[w], issue_free)
        for r in src: t = max(t, ready[w].get(r, 0.0))
        for r in dst: t = max(t, ready[w].get(r, 0.0) - lat(base, op) + 1)  # WAW approx
        if base in ('DFMA', 'DADD', 'DMUL', 'DSETP'): t = max(t, pipe_free)
        cand.append((t, w))
    t, w = min(cand)
    base, op, dst, src = P[pc[w]]
    if base in ('DFMA', 'DADD', 'DMUL', 'DSETP'): pipe_free = t + occ(base)
    issue_free = t + 1.0
    tnext[w] = t + 1.0
    for r in dst: ready[w][r] = t + lat(base, op)
    pc[w] += 1
    if pc[w] == len(P):
        pc[w] = 0; it[w] += 1
        if it[w] == 1: start[w] = t
        if it[w] == ITER: done[w] = t
per = sum((done[w] - start[w]) / (ITER - 1) for w in range(W)) / W
print(f'model: {per:.0f} cycles per row per warp with {W} warps; pipe floor {nD*2.4*W:.0f}; per-row/warps = {per/W:.0f}')
