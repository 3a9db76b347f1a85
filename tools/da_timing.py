#!/usr/bin/env python
"""Throughput of the sampler variants at the BASELINE sizes (chain-steps/s, tensor-core path): random-walk MH,
delayed acceptance (time-strided coarse stage) and component-wise MH."""
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

pb = helpers.load_problem("geometric")
C = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
x0 = np.tile(pb["x_opts"], (C // 5 + 1, 1))[:C]
sd = np.sqrt(pb["sigma2"] * np.diag(pb["cov"]))
for name, sampler, kw, scale0 in (("mh", E.SAMPLER_MH, {}, 0.128), ("da stride 4", E.SAMPLER_DA, {"da_stride": 4}, 0.128),
                                  ("da stride 2", E.SAMPLER_DA, {"da_stride": 2}, 0.128), ("cwmh", E.SAMPLER_CWMH, {"cw_scale": 1.2 * sd}, 0.1)):
    eng = helpers.engine(pb, C, 1)
    eng.set_sampler(sampler, **kw)
    eng.init_chains(x0, scale0=scale0, seed=3)
    eng.run(20, 0, 1, keep_samples=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.run(steps, 0, 1, keep_samples=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = eng.state()
    acc = st["accept_count"].double().mean().item() / (steps + 20)
    extra = ""
    if sampler == E.SAMPLER_DA:
        extra = f", promoted {eng.aux_state()['promote_count'].double().mean().item() / (steps + 20):.3f}"
    print(f"{name:12s} {C} chains: {C * steps / dt / 1e6:8.2f} M chain-steps/s, acceptance {acc:.3f}{extra}")
    eng.close()

# Q-factor problem (P = 4, sigma2 = 0.5): plain MH, strided DA and DA with the reference's FFT surrogate as first stage
pbq = helpers.load_problem("qfactor")
cf = helpers.fft_coarse()
x0q = np.tile(pbq["x_opts"], (C // 5 + 1, 1))[:C]
for name, setup in (("qf mh", None), ("qf da stride 4", "stride"), ("qf da fft", "fft"), ("qf da fft, coarse on CUDA cores", "fft_fp32")):
    os.environ.pop("MEMS_COARSE_FP32", None)
    if setup == "fft_fp32":                       # the same kernel with the coarse stage on the CUDA cores (round-1 arrangement)
        os.environ["MEMS_COARSE_FP32"] = "1"
        setup = "fft"
    eng = helpers.engine(pbq, C, 1)
    if setup == "stride":
        eng.set_sampler(E.SAMPLER_DA, da_stride=4)
    elif setup == "fft":
        eng.set_coarse_fft(cf["layers"], cf["scaler_mean"], cf["scaler_scale"])
        eng.set_sampler(E.SAMPLER_DA, da_stride=0)
    eng.init_chains(x0q, scale0=0.218, seed=3)
    eng.run(20, 0, 1, keep_samples=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.run(steps, 0, 1, keep_samples=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    acc = eng.state()["accept_count"].double().mean().item() / (steps + 20)
    extra = f", promoted {eng.aux_state()['promote_count'].double().mean().item() / (steps + 20):.3f}" if setup else ""
    print(f"{name:32s} {C} chains: {C * steps / dt / 1e6:8.2f} M chain-steps/s, acceptance {acc:.3f}{extra}")
    eng.close()
