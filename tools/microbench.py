#!/usr/bin/env python
"""Run the epilogue micro-benchmarks (mems_microbench) for several warp counts and print per-SM rates."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
from mcmc_mems_b200 import _lib  # noqa: E402

L = _lib.lib()
blocks, iters = 148, 400
# mode -> (name, bytes of TMEM traffic per warp-iteration or None, activations per thread-iteration or None)
MODES = {
    0: ("tcgen05.ld 32x32b.x32 + wait", 4096, None),
    4: ("2 x tcgen05.ld x32, one wait", 8192, None),
    5: ("tcgen05.ld 32x32b.x64 + wait", 8192, None),
    6: ("2 x tcgen05.ld 16x256b.x8, one wait", 8192, None),
    10: ("4 x tcgen05.ld x32 per wait", 4096, None),
    1: ("2 x tcgen05.st x16 + wait", 4096, None),
    2: ("tanh + split math", None, 32),
    7: ("tanh math only", None, 32),
    8: ("split + pack only", None, 32),
    9: ("tanh.approx + one pack", None, 32),
    3: ("ld + tanh + split + st", None, 32),
}
for mode, (name, nbytes, nact) in MODES.items():
    for warps in (4, 8, 16):
        cyc = torch.zeros(blocks, dtype=torch.int64, device="cuda")
        sink = torch.zeros(blocks * 32 * warps, dtype=torch.int32, device="cuda")
        for rep in range(2):
            _lib.check(L.mems_microbench(0, mode, warps, iters, blocks, C.c_void_p(cyc.data_ptr()), C.c_void_p(sink.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mems_microbench")
        torch.cuda.synchronize()
        c = cyc.double().median().item()
        per_iter_sm = c / iters                      # cycles for `warps` warp-iterations on one SM
        if nbytes:
            print(f"mode {mode:2d} {name:40s} warps {warps:2d}: {per_iter_sm:8.1f} cyc/iter/SM -> {warps * nbytes / per_iter_sm:7.1f} B/clk/SM")
        else:
            act = warps * 32 * nact
            print(f"mode {mode:2d} {name:40s} warps {warps:2d}: {per_iter_sm:8.1f} cyc/iter/SM -> {per_iter_sm / act * 8192:7.1f} cyc per 128x64 tile-layer")
