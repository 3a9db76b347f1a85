#!/usr/bin/env python
"""Posterior agreement, human-readable: the notebook workflow (5 starts x 6000 adaptive MH updates, burn 1000, thin 5;
tests/Xaccelerometer_geometric/identification.ipynb cells 4-16) through (a) the CPU oracle, (b) the GPU drop-in
(`MH(...).sample_adapt`), and (c) 4096 GPU chains, next to the numbers the reference notebook printed / stored."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from oracle import mh_sample_adapt  # noqa: E402
from mcmc_mems_b200 import (FixtureProcessor, NN_Model, Model, Gaussian, Uniform, JointDistribution, MH, Continuous1D,  # noqa: E402
                            Discrete, create_forward_model_function, diagnostics as D)

pb = helpers.load_problem("geometric")
pins = helpers.pins()
N, Nb, Nt = 6000, 1000, 5
P = 3


def row(label, mean, std, acc, scale, ess, rhat, secs):
    f = lambda a: "[" + " ".join(f"{v:10.6f}" for v in a) + "]"          # noqa: E731
    print(f"{label:34s} mean {f(mean)} std {f(std)} acc {acc:6.3f} scale {scale:7.4f} ess/1000 {np.round(ess).astype(int)} rhat {np.round(rhat, 4)} ({secs:6.2f} s)")


g = pins["geometric_sample_10"]
print("reference notebook: stored chain 0 (unthinned, unseeded y_obs): mean", np.round(g["mean"], 6), "std", np.round(g["std"], 6))
pr = pins["geometric_printed"]
print("reference notebook: printed acceptance", np.round(pr["acc"], 3), "scale", np.round(pr["scale"], 4), "rhat", pr["rhat"])

L = np.linalg.cholesky(pb["cov"])
t0 = time.perf_counter()
chains, accs, scales = [], [], []
for k in range(5):
    s, _, a, sc = mh_sample_adapt(helpers.oracle_logd(pb), pb["x_opts"][k], N, 0, rng=np.random.default_rng(k), chol=L)
    chains.append(s[:, Nb::Nt]); accs.append(a); scales.append(sc)
ch = np.stack(chains)                                                  # [5, P, 1000]
row("oracle (numpy, CUQIpy semantics)", ch.mean(axis=(0, 2)), ch.transpose(1, 0, 2).reshape(P, -1).std(axis=1), np.mean(accs), np.mean(scales),
    [np.mean([D.ess_bulk(ch[c, p][None]) for c in range(5)]) for p in range(P)], [D.rhat_rank(ch[:, p, :]) for p in range(P)], time.perf_counter() - t0)

dp = FixtureProcessor(pb["time"], pb["scaler_mean"], pb["scaler_scale"])
net = NN_Model(); net.set_layers(pb["layers"])
fwd = create_forward_model_function(dp, net)
model = Model(forward=fwd, range_geometry=Continuous1D(len(pb["time"])), domain_geometry=Discrete(["Overetch", "Offset", "Thickness"]))
x = Uniform(low=pb["low"], high=pb["high"])
y = Gaussian(mean=model(x), cov=pb["sigma2"] * np.eye(len(pb["time"])))
post = JointDistribution(x, y)(y=pb["y_obs"])
prop = Gaussian(mean=np.zeros(P), cov=pb["cov"])
for label, x0 in (("GPU drop-in, 5 chains", pb["x_opts"]), ("GPU drop-in, 4096 chains", np.tile(pb["x_opts"], (820, 1))[:4096])):
    t0 = time.perf_counter()
    b = MH(target=post, proposal=prop, x0=x0, seed=11).sample_adapt(N, 0).burnthin(Nb, Nt)
    secs = time.perf_counter() - t0
    S = b.samples                                                     # [C, P, 1000]
    sub = S[:: max(1, len(S) // 64)]                                   # diagnostics on a subset of chains (host arviz semantics)
    row(label, S.mean(axis=(0, 2)), S.transpose(1, 0, 2).reshape(P, -1).std(axis=1), float(np.mean(b.acc_rate)), float(np.mean(b.scale)),
        [np.mean([D.ess_bulk(sub[c, p][None]) for c in range(len(sub))]) for p in range(P)], [D.rhat_rank(sub[:, p, :]) for p in range(P)], secs)
