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


def _ll_of(pb, X, stride=1):
    """Oracle log-likelihood core of points X [n,P] on time points 0, stride, 2*stride, ... (rescaled by T/Tc),
    with -inf outside the prior box."""
    from oracle import batched_forward
    f = batched_forward(pb["time"], pb["scaler_mean"], pb["scaler_scale"], pb["layers"], X)
    T = f.shape[1]
    idx = np.arange(0, T, stride)
    r = pb["y_obs"][None, idx] - f[:, idx].astype(np.float64)
    ll = -0.5 * np.sum(r * r, axis=1) / pb["sigma2"] * (T / len(idx))
    inb = np.all((X >= pb["low"]) & (X <= pb["high"]), axis=1)
    return np.where(inb, ll, -np.inf)


def _tally(res, want, got, margin, band):
    if want != got:
        if margin <= band:
            res["near_ties"] += 1
        else:
            res["flips"] += 1


def check_explicit_cwmh(pb, x0, cw_scale, z, log_u, smp, llp, dec, band):
    """Component-wise MH replay (cuqi CWMH.single_update semantics).  z, log_u: [S,C,P];
    kernel outputs smp [S,P,C], llp/dec [S,P,C]."""
    from oracle import decide
    S, C, P = z.shape
    cur = np.asarray(x0, np.float64).copy()
    ll_cur = _ll_of(pb, cur)
    res = {"flips": 0, "near_ties": 0, "max_ll_err": 0.0, "state_err": 0.0}
    cw_scale = np.asarray(cw_scale, np.float64)
    for s in range(S):
        x_o = cur + (cw_scale[s] if cw_scale.ndim == 3 else cw_scale[None, :]) * z[s]     # [S,C,P]: per-step scales (adaptive runs)
        for p in range(P):
            prop = cur.copy()
            prop[:, p] = x_o[:, p]
            ll_star = _ll_of(pb, prop)
            k = llp[s, p]
            fin = np.isfinite(ll_star)
            assert np.array_equal(np.isfinite(k), fin)
            if fin.any():
                res["max_ll_err"] = max(res["max_ll_err"], float(np.max(np.abs(k[fin] - ll_star[fin]) / np.maximum(1.0, np.abs(ll_star[fin])))))
            for c in range(C):
                want, got = decide(log_u[s, c, p], ll_star[c], ll_cur[c]), bool(dec[s, p, c])
                _tally(res, want, got, abs(log_u[s, c, p] - min(0.0, ll_star[c] - ll_cur[c])), band)
                if got:
                    cur[c, p] = prop[c, p]
                    ll_cur[c] = ll_star[c]
        res["state_err"] = max(res["state_err"], float(np.max(np.abs(smp[s].T - cur))))
    return res


def fft_coarse():
    """Fixture of the FFT coarse surrogate (tests/golden/fft_coarse.npz): layers + the scaler of its inputs."""
    d = np.load(os.path.join(GOLDEN, "fft_coarse.npz"))
    return {"layers": layers_of(d), "scaler_mean": d["scaler_mean"], "scaler_scale": d["scaler_scale"],
            "X_test": d["X_test"], "y_test": d["y_test"], "y_series_test": d["y_series_test"]}


def coarse_ll_fn(pb, cf):
    """X [n,P] -> coarse log-likelihood core of the FFT surrogate on problem pb (oracle/coarse.py)."""
    from oracle.coarse import coarse_loglik
    return lambda X, _stride=None: coarse_loglik(cf["layers"], cf["scaler_mean"], cf["scaler_scale"], X, pb["y_obs"],
                                                 pb["sigma2"], pb["low"], pb["high"])


def check_explicit_da(pb, x0, scale, stride, xi, log_u, smp, llp, dec, band, coarse_ll=None):
    """Delayed-acceptance replay (Christen & Fox).  xi [S,C,P], log_u [S,C,2]; kernel llp/dec [S,2,C].
    coarse_ll: callable X -> coarse log-likelihood (default: the time-strided trace with `stride`)."""
    from oracle import decide
    S, C, P = xi.shape
    cur = np.asarray(x0, np.float64).copy()
    if coarse_ll is None:
        coarse_ll = lambda X: _ll_of(pb, X, stride)   # noqa: E731
    lf, lc = _ll_of(pb, cur), coarse_ll(cur)
    res = {"flips": 0, "near_ties": 0, "max_ll_err": 0.0, "state_err": 0.0, "promoted": 0, "accepted": 0}
    scale = np.asarray(scale, np.float64)
    for s in range(S):
        prop = cur + (scale[s][:, None] if scale.ndim == 2 else scale) * xi[s]             # [S,C]: per-step scales (adaptive runs)
        lc_star, lf_star = coarse_ll(prop), _ll_of(pb, prop)
        fin = np.isfinite(lc_star)
        assert np.array_equal(np.isfinite(llp[s, 0]), fin)
        if fin.any():
            res["max_ll_err"] = max(res["max_ll_err"], float(np.max(np.abs(llp[s, 0][fin] - lc_star[fin]) / np.maximum(1.0, np.abs(lc_star[fin])))))
        for c in range(C):
            want1, got1 = decide(log_u[s, c, 0], lc_star[c], lc[c]), bool(dec[s, 0, c])
            _tally(res, want1, got1, abs(log_u[s, c, 0] - min(0.0, lc_star[c] - lc[c])), band)
            got2 = bool(dec[s, 1, c])
            if got1:
                res["promoted"] += 1
                ratio = (lf_star[c] - lf[c]) - (lc_star[c] - lc[c])
                want2 = decide(log_u[s, c, 1], ratio, 0.0)
                _tally(res, want2, got2, abs(log_u[s, c, 1] - min(0.0, ratio)), band)
                if np.isfinite(lf_star[c]):
                    assert abs(llp[s, 1, c] - lf_star[c]) <= 1e-3 * max(1.0, abs(lf_star[c]))
                else:                       # promoted by the min(0, nan) quirk although outside the box: stays -inf
                    assert np.isneginf(llp[s, 1, c])
            else:
                assert not got2 and np.isneginf(llp[s, 1, c])
            if got2:
                res["accepted"] += 1
                cur[c] = prop[c]
                lf[c], lc[c] = lf_star[c], lc_star[c]
        res["state_err"] = max(res["state_err"], float(np.max(np.abs(smp[s].T - cur))))
    return res


def cuqi_scale_trajectory(dec, Na, scale0, target, first=1.0):
    """Scale used by every update of an adaptive run, recomputed from its decisions with the CUQIpy 0.8.0 rule
    (MH._sample_adapt / CWMH._sample_adapt): dec [S, ...] 0/1, acc[0] = `first` belongs to the first window; every Na
    updates lambda <- exp(log lambda + (mean(acc[idx:idx+Na]) - target)/sqrt(i+1)), scale = min(lambda, 1).
    Returns (scale_steps [S, ...], final scale [...])."""
    dec = np.asarray(dec, np.float64)
    S = dec.shape[0]
    acc = np.concatenate([np.full((1,) + dec.shape[1:], first), dec])
    lam = np.broadcast_to(np.asarray(scale0, np.float64), dec.shape[1:]).copy()
    scale = lam.copy()
    steps = np.empty_like(dec)
    i = idx = 0
    for s in range(S):
        steps[s] = scale
        if Na > 0 and (s + 1) % Na == 0:
            hat = acc[idx:idx + Na].mean(axis=0)
            lam = np.exp(np.log(lam) + (hat - target) / np.sqrt(i + 1))
            scale = np.minimum(lam, 1.0)
            i += 1
            idx += Na
    return steps, scale
