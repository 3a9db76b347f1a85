#!/usr/bin/env python
"""Where the watts go: board power and SM clock (nvidia-smi, sampled while the kernel runs) of the epilogue alone, the MMA
stream alone, and both together (micro-benchmark kernels of tools/libmems_b200_tools.so), next to the product kernel.
Evidence for DESIGN.md's statement that the bench launch is power-bound."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import helpers  # noqa: E402
from mcmc_mems_b200 import _lib  # noqa: E402

L = _lib.tools_lib()
blocks = 148
SECONDS = 4.0


def sampled(fn):
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.3)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < SECONDS:
        fn()
        torch.cuda.synchronize()
        n += 1
    clk = s.stop()
    pw = [float(r[3]) for r in s.rows if len(r) > 3]
    return clk, float(np.median(pw[3:])) if len(pw) > 6 else float("nan"), n


def micro(mode, warps, iters):
    cyc = torch.zeros(blocks, dtype=torch.int64, device="cuda")
    sink = torch.zeros(blocks * 544, dtype=torch.int32, device="cuda")

    def go():
        _lib.check_tools(L.mems_tools_microbench(0, mode, warps, iters, blocks, C.c_void_p(cyc.data_ptr()), C.c_void_p(sink.data_ptr()),
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mems_tools_microbench")
    go()
    torch.cuda.synchronize()
    clk, pw, n = sampled(go)
    return clk, pw, cyc.double().median().item() / iters


print(f"idle: {sampled(lambda: time.sleep(0.05))[1]:.0f} W")
for mode, warps, iters, name, per in ((200, 16, 4000, "epilogue alone, constant data (16 warps, no MMAs)", 2),
                                      (201, 16, 4000, "epilogue + MMA batches at the product's duty, constant data", 2),
                                      (210, 16, 4000, "epilogue alone, pseudo-random activations", 2),
                                      (211, 16, 4000, "epilogue + MMA batches at the product's duty, random data", 2),
                                      (212, 16, 4000, "epilogue + MMA batches back to back, random data", 2)):
    clk, pw, c = micro(mode, warps, iters)
    print(f"{name:62s}: {pw:6.0f} W  SM {clk['sm_mhz']:6.0f} MHz  {c / per:8.1f} cycles per {'MMA' if mode == 100 else 'tile-layer'}  {clk['reasons']}",
          flush=True)
pb = helpers.load_problem("geometric")
Cn = 1 << 20
eng = helpers.engine(pb, Cn, 1)
x0 = np.tile(pb["x_opts"], (Cn // 5 + 1, 1))[:Cn]
eng.init_chains(x0, scale0=0.128, seed=1)
eng.run(8, 0, 8, keep_samples=False)
torch.cuda.synchronize()
clk, pw, n = sampled(lambda: eng.run(16, 0, 16, keep_samples=False))
print(f"{'product kernel, sweep batch (1 048 576 chains x 16 updates)':62s}: {pw:6.0f} W  SM {clk['sm_mhz']:6.0f} MHz  {clk['reasons']}")
