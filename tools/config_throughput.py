#!/usr/bin/env python
"""Throughput of every BASELINE.json configuration on one GPU (tensor-core path): chain-steps/s, surrogate rows/s and
algorithmic TFLOP/s (2*((P+1)*64 + 5*64*64 + 64) FLOP per row, rows = chains * T * surrogate passes per update)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from mcmc_mems_b200 import engine as E  # noqa: E402


def z_problem(T=1000):
    from oracle import batched_forward
    pb = helpers.load_problem("geometric")
    pb["time"] = np.linspace(0.0, 1.49e-3, T)
    f = batched_forward(pb["time"], pb["scaler_mean"], pb["scaler_scale"], pb["layers"], pb["x_true"][None])[0]
    pb["y_obs"] = f + np.sqrt(pb["sigma2"]) * np.random.default_rng(0).standard_normal(T)
    return pb


def run(label, pb, C, steps, sampler=E.SAMPLER_MH, scale0=0.128, passes=1.0, **kw):
    P, T = len(pb["low"]), len(pb["time"])
    eng = helpers.engine(pb, C, 1)
    if kw.pop("fft", False):
        cf = helpers.fft_coarse()
        eng.set_coarse_fft(cf["layers"], cf["scaler_mean"], cf["scaler_scale"])
    eng.set_sampler(sampler, **kw)
    x0 = np.tile(pb["x_opts"], (C // 5 + 1, 1))[:C]
    eng.init_chains(x0, scale0=scale0, seed=3)
    eng.run(max(2, steps // 4), 0, 1, keep_samples=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.run(steps, 0, 1, keep_samples=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = eng.state()
    acc = st["accept_count"].double().mean().item() / (steps + max(2, steps // 4))
    if sampler == E.SAMPLER_DA:
        prom = eng.aux_state()["promote_count"].double().mean().item() / (steps + max(2, steps // 4))
        passes = (0.0 if kw.get("da_stride", 1) == 0 else 1.0 / kw["da_stride"]) + prom     # fine passes actually evaluated
    rows = C * steps * T * passes / dt
    flop_row = 2 * ((P + 1) * 64 + 5 * 64 * 64 + 64)
    print(f"{label:58s} {C * steps / dt / 1e6:8.2f} M chain-steps/s  {rows / 1e9:6.2f} G rows/s  {rows * flop_row / 1e12:6.1f} TFLOP/s  acc {acc:.3f}")
    eng.close()


geo, qf = helpers.load_problem("geometric"), helpers.load_problem("qfactor")
sd = np.sqrt(geo["sigma2"] * np.diag(geo["cov"]))
cov1 = np.zeros((4, 4)); cov1[3, 3] = qf["cov"][3, 3]
run("configs[0] 1-parameter Q posterior, 1 chain", dict(qf, cov=cov1), 1, 2000, scale0=0.5)
run("configs[1] geometric, 4096 chains (bench workload)", geo, 4096, 256)
run("configs[2] Z trace T=1000, 65536 chains, component-wise MH", z_problem(), 65536, 6, sampler=E.SAMPLER_CWMH, scale0=0.1, passes=3.0, cw_scale=0.3 * sd)
run("configs[3] delayed acceptance (stride 4), 262144 chains", geo, 262144, 40, sampler=E.SAMPLER_DA, da_stride=4)
run("configs[3'] delayed acceptance (FFT stage), Q-factor, 262144", qf, 262144, 40, sampler=E.SAMPLER_DA, scale0=0.218, da_stride=0, fft=True)
run("configs[4] sweep, 1048576 chains", geo, 1048576, 20)
