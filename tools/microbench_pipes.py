#!/usr/bin/env python
"""Round-2 pipe probes: which instruction classes of the epilogue share an execution pipe, and what the epilogue's
instruction mix costs without its dependencies (mems_tools.cu pipe_probe_kernel modes 32..47)."""
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
blocks, iters = 148, 2000
PROBES = {
    23: ("FFMA2 alone", 16), 24: ("LOP3 alone", 16), 21: ("F2FP alone", 16), 35: ("FFMA (scalar, 3-reg) alone", 16),
    46: ("PRMT alone", 16), 47: ("HADD2 alone", 16),
    32: ("LOP3 + FFMA2 pairs", 16), 33: ("F2FP + FFMA2 pairs", 16), 34: ("F2FP + LOP3 pairs", 16), 36: ("FFMA + LOP3 pairs", 16),
    45: ("FFMA2 + FFMA pairs", 16),
    48: ("FHADD (add.f32.f16) alone", 16), 49: ("FHADD + FFMA2 pairs", 16), 54: ("FHADD + LOP3 pairs", 16),
    55: ("round-2 epilogue mix (per activation)", 16), 56: ("round-2 mix without the sign LOP3", 16),
    37: ("epilogue mix, all classes (per activation)", 16), 38: ("mix without LOP3", 16), 39: ("mix without F2FP", 16),
    40: ("mix without MUFU", 16), 41: ("mix without packed FP32", 16), 42: ("mix with half the LOP3", 16),
    43: ("LOP3 + F2FP part only", 16), 44: ("MUFU + packed FP32 part only", 16),
}
if len(sys.argv) > 1:
    PROBES = {int(m): PROBES[int(m)] for m in sys.argv[1].split(",")}
for mode, (name, nops) in PROBES.items():
    for warps in (4, 8, 16):
        cyc = torch.zeros(blocks, dtype=torch.int64, device="cuda")
        sink = torch.zeros(blocks * 32 * warps, dtype=torch.int32, device="cuda")
        for rep in range(2):
            _lib.check_tools(L.mems_tools_microbench(0, mode, warps, iters, blocks, C.c_void_p(cyc.data_ptr()), C.c_void_p(sink.data_ptr()),
                                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mems_tools_microbench")
        torch.cuda.synchronize()
        c = cyc.double().median().item() / iters
        print(f"probe {mode} {name:46s} warps {warps:2d}: {c / nops / (warps / 4):6.2f} SMSP-cycles per warp-group "
              f"(x64 = {c / nops / (warps / 4) * 64:6.0f} per tile-layer)", flush=True)
