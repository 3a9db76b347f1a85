#!/usr/bin/env python
"""Wall time of the reference-shaped call MH(target, proposal, x0).sample_adapt(6000, 0) through the drop-in at small chain
counts (the reference's own size is 5 chains x 6000 updates, identification.ipynb:298-302): first call (CUDA context, engine,
weight image) and a repeat, next to the numpy oracle on one host core."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcmc_mems_b200 import Gaussian, MH, fixtures  # noqa: E402

fx = fixtures.load_fixture("geometric")
post = fixtures.posterior_of(fx)
prop = Gaussian(mean=np.zeros(3), cov=fx.cov)
N = 6000
torch.zeros(1, device="cuda")
for C in (1, 5, 64, 512, 4096):
    x0 = np.tile(fx.x_opts, (C // 5 + 1, 1))[:C]
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        b = MH(target=post, proposal=prop, x0=x0 if C > 1 else x0[0], seed=rep).sample_adapt(N, 0)
        s = b.samples
        ts.append(time.perf_counter() - t0)
    print(f"{C:5d} chains x {N} adaptive updates: first call {ts[0]:6.3f} s, repeat {min(ts[1:]):6.3f} s "
          f"= {C * N / min(ts[1:]) / 1e6:8.4f} M chain-steps/s", flush=True)
from oracle import mh_sample_adapt  # noqa: E402
import bench  # noqa: E402
logd = bench.oracle_logd_of(fx)
t0 = time.perf_counter()
mh_sample_adapt(logd, fx.x_opts[0], N, 0, rng=np.random.default_rng(0), chol=np.linalg.cholesky(fx.cov))
dt = time.perf_counter() - t0
print(f"numpy oracle, one chain x {N} updates on one host core: {dt:6.3f} s = {N / dt / 1e3:.2f} k chain-steps/s")
