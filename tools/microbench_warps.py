#!/usr/bin/env python
"""Round-2 question: does the stand-alone epilogue keep speeding up past 16 warps per SM?  Runs the epilogue bodies
(mems_tools.cu modes 11, 15, 28, 30, 31) at 16..32 warps, each compiled under the register budget that warp count leaves
(mode + 1000: 80 registers / 24 warps, mode + 2000: 64 registers / 32 warps)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
from mcmc_mems_b200 import _lib  # noqa: E402

L = _lib.tools_lib()
blocks, iters = 148, 400
NAMES = {67: "4 x ld8 with prefetch", 68: "2 x ld16, rcp/16, prefetch, one wait::st", 60: "4 x ld8, 1/8 poly exp2", 61: "4 x ld8, 1/4 poly exp2", 62: "2 x ld16, rcp per 16", 63: "2 x ld16, rcp/16, 1/8 poly",
         64: "2 x ld16, rcp/16, 1/4 poly", 65: "2 x ld16, rcp/8, 1/4 poly", 66: "2 x ld16, rcp/16, 1/2 poly", 50: "ld32 + tanh + split + st, round-2 arithmetic", 51: "2 x (ld16 ... st8x2 + wait), round-2 arithmetic",
         52: "pipelined across chunks, round-2 arithmetic", 53: "4 x (ld8 ... st4x2), round-2 arithmetic", 11: "ld32 + tanh + split + st (plain)", 15: "pipelined tanh(next) || split+st(cur)", 28: "pipelined, 16 cols per iteration",
         30: "2 x (ld16 + tanh + split + st8x2 + wait)", 31: "2 x ld16, one wait::ld / one wait::st"}
NACT = {67: 32, 68: 32, 60: 32, 61: 32, 62: 32, 63: 32, 64: 32, 65: 32, 66: 32, 11: 32, 15: 32, 28: 16, 30: 32, 31: 32, 50: 32, 51: 32, 52: 32, 53: 32}
MODES = [int(m) for m in sys.argv[1].split(",")] if len(sys.argv) > 1 else (11, 15, 28, 30, 31, 50, 51, 52, 53)
for mode in MODES:
    for warps in ((16, 24, 32) if len(sys.argv) > 2 else (8, 16, 20, 24, 28, 32)):
        variants = [(0, 128)] if warps <= 16 else ([(1000, 80)] if warps <= 24 else []) + [(2000, 64)]
        if warps in (16, 24):
            variants = [(0, 128), (1000, 80), (2000, 64)] if warps == 16 else [(1000, 80), (2000, 64)]
        for off, regs in variants:
            cyc = torch.zeros(blocks, dtype=torch.int64, device="cuda")
            sink = torch.zeros(blocks * 32 * warps, dtype=torch.int32, device="cuda")
            for rep in range(2):
                _lib.check_tools(L.mems_tools_microbench(0, mode + off, warps, iters, blocks, C.c_void_p(cyc.data_ptr()), C.c_void_p(sink.data_ptr()),
                                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mems_tools_microbench")
            torch.cuda.synchronize()
            c = cyc.double().median().item() / iters
            act = warps * 32 * NACT[mode]
            print(f"mode {mode:2d} {NAMES[mode]:42s} warps {warps:2d} regs<={regs:3d}: {c:8.1f} cyc/iter/SM -> {c / act * 8192:7.1f} cyc per 128x64 tile-layer", flush=True)
