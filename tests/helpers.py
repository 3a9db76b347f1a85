"""Shared test helpers: golden fixtures and the oracle wiring."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = "/root/reference"


def load_problem(name):
    """name in {'geometric', 'qfactor'} -> dict with layers, time, scaler constants, y_obs, ..."""
    d = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    out = {k: d[k] for k in d.files}
    out["layers"] = layers_of(d)
    out["sigma2"] = float(d["sigma2"])
    return out


def layers_of(d):
    return [(d[f"W{i}"], d[f"b{i}"], str(d["activations"][i])) for i in range(int(d["n_layers"]))]


def pins():
    return json.load(open(os.path.join(GOLDEN, "pins.json")))


def oracle_forward(pb):
    from oracle import make_forward_model
    return make_forward_model(pb["time"], pb["scaler_mean"], pb["scaler_scale"], pb["layers"])


def oracle_logd(pb):
    from oracle import make_posterior_logd
    return make_posterior_logd(oracle_forward(pb), pb["y_obs"], pb["sigma2"], pb["low"], pb["high"])


def loglik_core(pb, f):
    """The device convention: -0.5*sum((y_obs-f)^2)/sigma2 (no additive constants)."""
    r = pb["y_obs"][None, :] - np.asarray(f, np.float64)
    return -0.5 * np.sum(r * r, axis=-1) / pb["sigma2"]


def loglik_const(pb):
    T = len(pb["time"])
    return -0.5 * T * (np.log(pb["sigma2"]) + np.log(2 * np.pi)) + np.log(1.0 / np.prod(pb["high"] - pb["low"]))


def spec_of(pb):
    from mcmc_mems_b200 import SurrogateSpec
    return SurrogateSpec(pb["layers"], pb["time"], pb["scaler_mean"], pb["scaler_scale"])


def engine(pb, n_chains, path, **kw):
    from mcmc_mems_b200 import MHEngine
    return MHEngine(spec_of(pb), pb["y_obs"], pb["sigma2"], pb["low"], pb["high"], pb["cov"],
                    n_chains=n_chains, path=path, **kw)


def check_explicit_run(pb, x0, scale, xi, log_u, smp, llp, dec, band, ll0=None):
    """Stage-isolated parity check of an explicit-randomness run against the oracle.

    x0 [C,P]; xi [S,C,P]; log_u [S,C]; kernel outputs smp [S,P,C], llp [S,C], dec [S,C]
    (numpy).  For every update the proposal is rebuilt from the kernel's own previous state
    (so one legitimate near-tie flip cannot snowball), the oracle evaluates it, and the oracle's
    decision rule is applied.  Returns dict(flips, near_ties, max_ll_err, state_err).
    """
    from oracle import batched_forward, decide
    S, C, P = xi.shape
    prev = np.asarray(x0, np.float64).copy()                       # [C,P]
    f0 = batched_forward(pb["time"], pb["scaler_mean"], pb["scaler_scale"], pb["layers"], prev)
    ll_cur = loglik_core(pb, f0)
    inb0 = np.all((prev >= pb["low"]) & (prev <= pb["high"]), axis=1)
    ll_cur = np.where(inb0, ll_cur, -np.inf)
    flips = near = 0
    max_err = 0.0
    state_err = 0.0
    scale = np.broadcast_to(np.asarray(scale, np.float64), (S, C))
    for s in range(S):
        prop = prev + scale[s][:, None] * xi[s]
        f = batched_forward(pb["time"], pb["scaler_mean"], pb["scaler_scale"], pb["layers"], prop)
        ll_star = loglik_core(pb, f)
        inb = np.all((prop >= pb["low"]) & (prop <= pb["high"]), axis=1)
        ll_star = np.where(inb, ll_star, -np.inf)
        k = llp[s]
        assert np.array_equal(np.isfinite(k), inb), "prior-box decisions differ"
        if inb.any():
            max_err = max(max_err, float(np.max(np.abs(k[inb] - ll_star[inb]) / np.maximum(1.0, np.abs(ll_star[inb])))))
        for c in range(C):
            want = decide(log_u[s, c], ll_star[c], ll_cur[c])
            got = bool(dec[s, c])
            if want != got:
                margin = abs(log_u[s, c] - min(0.0, ll_star[c] - ll_cur[c]))
                if margin <= band:
                    near += 1
                else:
                    flips += 1
            # follow the KERNEL's decision so the next proposal is rebuilt from its state
            if got:
                prev[c] = prop[c]
                ll_cur[c] = ll_star[c]
        state_err = max(state_err, float(np.max(np.abs(smp[s].T - prev))))
    return {"flips": flips, "near_ties": near, "max_ll_err": max_err, "state_err": state_err}
