#!/usr/bin/env python
"""Timeline of the hand-offs inside ONE CTA of tc_run_kernel (development build with -DMEMS_TRACE=1, selected with MEMS_LIB):
where a hidden warp's time goes (waiting for the accumulator / epilogue body / between), how long an MMA batch takes from
commit to the first waiter, how far apart the warps of a group arrive.
    python tools/ab_bench.py --build-only trace=MEMS_TRACE=1
    MEMS_LIB=$PWD/mcmc-mems_b200/libmems_b200_trace.so python tools/trace_timeline.py"""
import ctypes as C
import os
import sys
from collections import defaultdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from mcmc_mems_b200 import _lib  # noqa: E402

pb = helpers.load_problem("geometric")
Cn = 148 * 2 * 128 * 2                       # two batches of 128 chains per group
eng = helpers.engine(pb, Cn, 1)
x0 = np.tile(pb["x_opts"], (Cn // 5 + 1, 1))[:Cn]
eng.init_chains(x0, scale0=0.128, seed=1)
eng.run(4, 0, 4, keep_samples=False)
torch.cuda.synchronize()
L = C.CDLL(_lib.LIB_PATH)
buf = np.zeros((32, 8192), np.uint64)
cnt = np.zeros(32, np.uint32)
L.mems_trace_read(buf.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))      # warm-up launch: discard
eng.run(2, 0, 2, keep_samples=False)
torch.cuda.synchronize()
L.mems_trace_read(buf.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
ev = {}
for w in range(32):
    n = int(cnt[w])
    t = (buf[w, :n] >> np.uint64(12)).astype(np.int64)
    tag = (buf[w, :n] & np.uint64(0xFFF)).astype(np.int64)
    ev[w] = [(int(tt), int(k & 15), int((k >> 4) & 3), int(k >> 6)) for tt, k in zip(t, tag)]     # (clock, kind, slot, layer)
NG = int(os.environ.get("TRACE_NGROUP", "2"))            # groups of the traced build (MEMS_NGROUP)
HW = 16 // NG                                             # hidden warps per group
NTAILW = 4 * NG
print("events per warp:", [len(ev[w]) for w in range(16 + NTAILW + 4)])
# hidden warps 0..15: group = w // 8
tot = defaultdict(lambda: np.zeros(3))
arrive = defaultdict(dict)                    # (group, occurrence of (slot, layer)) -> warp -> clock of the arrival
got = defaultdict(dict)
for w in range(16):
    e = ev[w]
    occ = defaultdict(int)
    last3 = None
    for i, (t, k, s, l) in enumerate(e):
        if k == 1:
            if last3 is not None:
                tot[w][2] += t - last3
            t1 = t
        elif k == 2:
            tot[w][0] += t - t1
            t2 = t
            key = (w // HW, s, l, occ[(s, l, 2)])
            occ[(s, l, 2)] += 1
            got[key][w] = t
        elif k == 3:
            tot[w][1] += t - t2
            last3 = t
            key = (w // HW, s, l, occ[(s, l, 3)])
            occ[(s, l, 3)] += 1
            arrive[key][w] = t
n_tl = sum(1 for (t, k, s, l) in ev[0] if k == 3)
print(f"hidden warp 0: {n_tl} tile-layers traced")
for w in range(16):
    a = tot[w] / max(n_tl, 1)
    print(f"  hidden warp {w:2d} (group {w // HW}, quarter {w % 4}, part {(w // 4) % (HW // 4)}): wait {a[0]:7.0f}  body {a[1]:7.0f}  between {a[2]:6.0f}  "
          f"= {a.sum():7.0f} cycles per tile-layer of its group")
# spread of arrivals within a group and MMA latency (issuer commit -> first hidden warp past the wait)
spread = [max(d.values()) - min(d.values()) for d in arrive.values() if len(d) == HW]
print(f"arrival spread within a group (last - first of its hidden warps): mean {np.mean(spread):.0f}  p50 {np.percentile(spread, 50):.0f}  "
      f"p90 {np.percentile(spread, 90):.0f} cycles")
iss = {}
IW0 = 16 + NTAILW
for w in range(IW0, IW0 + 4):
    g, s0 = (w - IW0) // (4 // NG), (w - IW0) % (4 // NG)
    occ = defaultdict(int)
    waits, issues = [], []
    for (t, k, s, l) in ev[w]:
        if k == 4:
            t4 = t
        elif k == 5:
            waits.append(t - t4)
            t5 = t
        elif k == 6:
            issues.append(t - t5)
            iss[(g, s, l, occ[(s, l)])] = t
            occ[(s, l)] += 1
    if waits:
        print(f"  issuer warp {w} (group {g}, slot {s0}): waits for operands {np.mean(waits):6.0f}, issue of a batch {np.mean(issues):5.0f} cycles")
lat_first, lat_last, gap = [], [], []
for key, t6 in iss.items():
    g, s, l, o = key
    if key in got and len(got[key]) == HW:
        lat_first.append(min(got[key].values()) - t6)
        lat_last.append(max(got[key].values()) - t6)
    prev = (g, s, l - 1, o) if l > 0 else None
    if prev and prev in arrive and len(arrive[prev]) == HW:
        gap.append(t6 - max(arrive[prev].values()))
print(f"commit -> first hidden warp sees the accumulator: mean {np.mean(lat_first):.0f} p90 {np.percentile(lat_first, 90):.0f}; "
      f"-> last: mean {np.mean(lat_last):.0f} cycles")
print(f"last arrival of a layer's operand -> MMA batch issued and committed: mean {np.mean(gap):.0f} p90 {np.percentile(gap, 90):.0f} cycles")
for w in range(16, 16 + NTAILW):
    e = ev[w]
    waits = [e[i + 1][0] - e[i][0] for i in range(len(e) - 1) if e[i][1] == 7 and e[i + 1][1] == 8]
    work = [e[i + 1][0] - e[i][0] for i in range(len(e) - 1) if e[i][1] == 8 and e[i + 1][1] == 9]
    if w in (16, 20):
        print(f"  tail warp {w}: waits for a tile's output {np.mean(waits):7.0f}, residual + next layer-1 operand {np.mean(work):5.0f} cycles")
