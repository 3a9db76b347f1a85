#!/usr/bin/env python
"""One micro-benchmark launch (for ncu captures): python tools/microbench_one.py MODE WARPS [ITERS]."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcmc_mems_b200 import _lib  # noqa: E402

mode, warps = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 400
L = _lib.tools_lib()
blocks = 148
cyc = torch.zeros(blocks, dtype=torch.int64, device="cuda")
sink = torch.zeros(blocks * 32 * warps, dtype=torch.int32, device="cuda")
for rep in range(2):
    _lib.check_tools(L.mems_tools_microbench(0, mode, warps, iters, blocks, C.c_void_p(cyc.data_ptr()), C.c_void_p(sink.data_ptr()),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mems_tools_microbench")
torch.cuda.synchronize()
print(mode, warps, cyc.double().median().item() / iters)
