#!/usr/bin/env python
"""Print (not assert) the numerical behaviour of the tensor-core plumbing on the GPU: accumulate
precision of tcgen05.mma kind::f16, accuracy of the split chain, forward error of both paths."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

g.build()
import helpers  # noqa: E402
from mcmc_mems_b200 import _lib  # noqa: E402

L = _lib.lib()


def selftest(variant, A, B):
    A = A.half().cuda().contiguous()
    B = B.half().cuda().contiguous()
    D = torch.full((128, 64), float("nan"), device="cuda")
    _lib.check_tools(_lib.tools_lib().mems_tools_selftest_umma(0, variant, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()),
                                    ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "selftest")
    torch.cuda.synchronize()
    return D.double().cpu()


gen = torch.Generator().manual_seed(0)
for v in (0, 1):
    A = torch.randn(128, 64, generator=gen).half()
    B = torch.randn(64, 64, generator=gen).half()
    D = selftest(v, A, B)
    want = A.double() @ B.double().t()
    print(f"variant {v}: max abs err {float((D - want).abs().max()):.3e}  (|D| ~ {float(want.abs().mean()):.2f}, "
          f"fp32-sum level would be ~{float(want.abs().mean()) * 2 ** -23:.1e})")

# bit depth of the accumulation: D = 64 + 64 * 2^-e via separate lo*hi MMAs (as in the product kernel)
print("separate-instruction small terms (variant 2): e, got-64, want")
for e in (8, 11, 12, 14, 16, 18, 20, 22, 23, 24):
    Ahi = torch.ones(128, 64)
    Alo = torch.full((128, 64), 2.0 ** -e)
    Whi = torch.ones(64, 64)
    Wlo = torch.zeros(64, 64)
    D = selftest(2, torch.cat([Ahi, Alo]), torch.cat([Whi, Wlo]))
    print(f"   e={e:2d}  got-64={float(D[0, 0]) - 64:.10e}  want={64 * 2.0 ** -e:.10e}")
print("same-instruction small terms (variant 0): K 0..7 = 1, K 8..15 = 2^-e, rest 0 -> 8 + 8*2^-e")
for e in (8, 11, 12, 14, 16, 18, 20, 22, 23, 24):
    A = torch.zeros(128, 64)
    A[:, :8] = 1.0
    A[:, 8:16] = 2.0 ** -e
    B = torch.ones(64, 64)
    D = selftest(0, A, B)
    print(f"   e={e:2d}  got-8={float(D[0, 0]) - 8:.10e}  want={8 * 2.0 ** -e:.10e}")

# split chain accuracy
A = (torch.rand(128, 64, generator=gen) * 2 - 1)
W = torch.randn(64, 64, generator=gen) * 1.5
Ahi = (A.view(torch.int32) & -8192).view(torch.float32)
Alo = (A - Ahi)
Whi = W.half().float()
Wlo = (W - Whi)
D = selftest(2, torch.cat([Ahi, Alo]), torch.cat([Whi, Wlo]))
want = A.double() @ W.double().t()
emu = (Ahi.double() @ Whi.double().t() + Alo.half().double() @ Whi.double().t() + Ahi.double() @ Wlo.half().double().t())
print(f"split chain: max abs err vs fp64 {float((D - want).abs().max()):.3e}; vs exact-arithmetic emulation "
      f"{float((D - emu).abs().max()):.3e}; emulation vs fp64 {float((emu - want).abs().max()):.3e}; |D|~{float(want.abs().mean()):.2f}")

# forward error of both paths
from oracle import batched_forward  # noqa: E402
pb = helpers.load_problem("geometric")
rng = np.random.default_rng(1)
Lc = np.linalg.cholesky(pb["cov"] * pb["sigma2"])
X = np.vstack([rng.uniform(pb["low"], pb["high"], size=(200, 3)), pb["x_opts"][0] + rng.standard_normal((200, 3)) @ Lc.T])
want = batched_forward(pb["time"], pb["scaler_mean"], pb["scaler_scale"], pb["layers"], X, np.float64)
for path, nm in ((0, "fp32"), (1, "tc")):
    try:
        eng = helpers.engine(pb, 1, path)
        y, ll = eng.forward(X)
        y = y.cpu().numpy().astype(np.float64)
        err = np.abs(y - want)
        print(f"forward {nm}: max abs {err.max():.3e} rms {np.sqrt((err ** 2).mean()):.3e} mean signed {np.mean(y - want):+.3e} "
              f"(box {err[:200].max():.3e}, posterior {err[200:].max():.3e})")
        llw = helpers.loglik_core(pb, want)
        d = ll.cpu().numpy() - llw
        print(f"   loglik: max rel {np.max(np.abs(d) / np.maximum(1, np.abs(llw))):.3e}; posterior delta-ll err rms "
              f"{np.std(np.diff(d[200:])):.4f} max {np.max(np.abs(np.diff(d[200:]))):.4f}")
        eng.close()
    except Exception as ex:  # keep probing
        print(f"forward {nm}: FAILED {ex}")
