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

L = _lib.tools_lib()
blocks, iters = 148, 400
if os.environ.get("MB_STAGGER"):
    iters = 401                # odd: warps of an SMSP start a quarter iteration apart (de-phased)
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
    11: ("ld + tanh(mix 0) + split + st", None, 32),
    12: ("ld + tanh(mix 1/4 poly) + split + st", None, 32),
    13: ("ld + tanh(mix 3/8 poly) + split + st", None, 32),
    14: ("ld + tanh(mix 1/2 poly) + split + st", None, 32),
    15: ("pipelined: tanh(next) || split+st(cur)", None, 32),
    16: ("pipelined, 1/4 poly", None, 32),
    17: ("64 cols/thread, intra-thread pipeline", None, 64),
    28: ("pipelined, 16 cols per iteration", None, 16),
    29: ("pipelined, st.x4 per group", None, 32),
    18: ("32 cols, rolled 8-col pipeline", None, 32),
    19: ("64 cols, rolled 8-col pipeline", None, 64),
}
if len(sys.argv) > 1:
    MODES = {int(m): MODES[int(m)] for m in sys.argv[1].split(",") if int(m) in MODES}
for mode, (name, nbytes, nact) in MODES.items():
    for warps in (4, 8, 12, 16):
        cyc = torch.zeros(blocks, dtype=torch.int64, device="cuda")
        sink = torch.zeros(blocks * 32 * warps, dtype=torch.int32, device="cuda")
        for rep in range(2):
            _lib.check_tools(L.mems_tools_microbench(0, mode, warps, iters, blocks, C.c_void_p(cyc.data_ptr()), C.c_void_p(sink.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mems_tools_microbench")
        torch.cuda.synchronize()
        c = cyc.double().median().item()
        per_iter_sm = c / iters                      # cycles for `warps` warp-iterations on one SM
        if nbytes:
            print(f"mode {mode:2d} {name:40s} warps {warps:2d}: {per_iter_sm:8.1f} cyc/iter/SM -> {warps * nbytes / per_iter_sm:7.1f} B/clk/SM")
        else:
            act = warps * 32 * nact
            print(f"mode {mode:2d} {name:40s} warps {warps:2d}: {per_iter_sm:8.1f} cyc/iter/SM -> {per_iter_sm / act * 8192:7.1f} cyc per 128x64 tile-layer")

# pipe throughput probes: 16 ops (mode 22: 16+16, mode 25: 16+64) per thread-iteration
PROBES = {20: ("MUFU.EX2", 16), 26: ("MUFU.RCP", 16), 21: ("F2FP cvt.rn.f16x2.f32", 16), 22: ("MUFU.EX2 + F2FP pairs", 16),
          23: ("FFMA2", 16), 27: ("FADD2", 16), 24: ("LOP3", 16), 25: ("MUFU.EX2 + 4 FFMA2", 16)}
if len(sys.argv) > 1:
    PROBES = {int(m): PROBES[int(m)] for m in sys.argv[1].split(",") if int(m) in PROBES}
for mode, (name, nops) in PROBES.items():
    for warps in (4, 8, 16):
        cyc = torch.zeros(blocks, dtype=torch.int64, device="cuda")
        sink = torch.zeros(blocks * 32 * warps, dtype=torch.int32, device="cuda")
        for rep in range(2):
            _lib.check_tools(L.mems_tools_microbench(0, mode, warps, 2000, blocks, C.c_void_p(cyc.data_ptr()), C.c_void_p(sink.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mems_tools_microbench")
        torch.cuda.synchronize()
        c = cyc.double().median().item() / 2000
        print(f"probe {mode} {name:24s} warps {warps:2d}: {c / nops / (warps / 4):6.2f} SMSP-cycles per warp-op (group)")

# tensor-core issue rate: M=128, K=16, kind::f16, N = 64 / 128 / 256, A operand in TMEM or shared memory
for mode, where in ((100, "TMEM"), (101, "smem")) if len(sys.argv) <= 1 else ():
    for n in (64, 128, 256):
        cyc = torch.zeros(blocks, dtype=torch.int64, device="cuda")
        sink = torch.zeros(64, dtype=torch.int32, device="cuda")
        it = 256
        for rep in range(2):
            _lib.check_tools(L.mems_tools_microbench(0, mode, n // 8, it, blocks, C.c_void_p(cyc.data_ptr()), C.c_void_p(sink.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mems_tools_microbench")
        torch.cuda.synchronize()
        c = cyc.double().median().item() / (4 * it)
        print(f"mma M=128 N={n:3d} K=16 A in {where}: {c:6.1f} cyc/MMA -> {2 * 128 * n * 16 / c:7.0f} FLOP/clk/SM")
