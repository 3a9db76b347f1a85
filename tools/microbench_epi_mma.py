#!/usr/bin/env python
"""Does concurrent tensor-core work slow the epilogue warps down?  16 epilogue warps (round-2 body) with no MMAs, with the
product kernel's MMA duty, and with back-to-back MMA batches on the same TMEM slots (mems_tools.cu epi_mma_kernel)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcmc_mems_b200 import _lib  # noqa: E402

L = _lib.tools_lib()
blocks, iters = 148, 400
for mode, name in ((200, "no MMAs"), (201, "one batch of 13 MMAs, then ~1100 cycles pause"), (202, "MMA batches back to back")):
    cyc = torch.zeros(blocks, dtype=torch.int64, device="cuda")
    sink = torch.zeros(blocks * 544, dtype=torch.int32, device="cuda")
    for rep in range(2):
        _lib.check_tools(L.mems_tools_microbench(0, mode, 16, iters, blocks, C.c_void_p(cyc.data_ptr()), C.c_void_p(sink.data_ptr()),
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mems_tools_microbench")
    torch.cuda.synchronize()
    c = cyc.double().median().item() / iters
    print(f"epilogue, 16 warps, {name:48s}: {c / 2:7.1f} cyc per 128x64 tile-layer", flush=True)
