"""Oracle: CUQIpy-0.8.0-semantics samplers and distributions (numpy).  Test infrastructure only.

The arithmetic restated here lives in the un-vendored dependency ``cuqipy==0.8.0``
(reference README.md:36).  Call sites in the reference that fix the semantics:
``tests/Xaccelerometer_geometric/identification.ipynb:161,165,169,292,300-302,308``
(cells 10 and 14) and the same cells of
``tests/Xaccelerometer_quality_factor/identification.ipynb:378-386,516-532``.

Randomness is injectable: every sampler accepts explicit ``xi`` (already
correlated proposal increments, one row per step) and ``log_u`` arrays so that the
CUDA path can be compared decision by decision on identical inputs.
"""
from __future__ import annotations

import math

import numpy as np


# --------------------------------------------------------------------------
# distributions (cuqi.distribution.Gaussian / Uniform / Posterior, as used)
# --------------------------------------------------------------------------
def gaussian_loglike(y_obs, f, sigma2):
    """``Gaussian(mean=f, cov=sigma2*I).logd(y_obs)`` (identification.ipynb:169).

    ``-0.5*(T*log(sigma2) + sum((y_obs-f)^2)/sigma2 + T*log(2*pi))`` with the
    residual formed in float64 (``y_obs`` is float64, ``f`` the float32 model output).
    """
    y = np.asarray(y_obs, np.float64)
    r = y - np.asarray(f, np.float64)
    T = y.shape[-1]
    return -0.5 * (T * math.log(sigma2) + np.sum(r * r, axis=-1) / sigma2 + T * math.log(2.0 * math.pi))


def uniform_logpdf(x, low, high):
    """``Uniform(low, high).logpdf(x)`` (identification.ipynb:165): -inf if any
    coordinate is strictly outside ``[low, high]``, else ``log(1/prod(high-low))``."""
    x = np.asarray(x, np.float64)
    low = np.asarray(low, np.float64)
    high = np.asarray(high, np.float64)
    if np.any(x < low) or np.any(x > high):
        return -np.inf
    return math.log(1.0 / np.prod(high - low))


def make_posterior_logd(forward, y_obs, sigma2, low, high):
    """``JointDistribution(x, y)(y=y_obs).logd`` (identification.ipynb:292):
    likelihood + prior, both always evaluated."""
    def logd(x):
        return gaussian_loglike(y_obs, forward(x), sigma2) + uniform_logpdf(x, low, high)
    return logd


def decide(log_u, target_star, target_cur):
    """CUQIpy ``MH.single_update`` acceptance test: ``log_u <= min(0, l* - l)``.

    Python's ``min(0, ratio)`` returns 0 when ``ratio`` is NaN (``-inf - -inf``),
    which is reproduced here explicitly.
    """
    ratio = target_star - target_cur
    alpha = ratio if ratio < 0 else 0.0
    return bool(log_u <= alpha)


# --------------------------------------------------------------------------
# random-walk MH (cuqi.sampler.MH)
# --------------------------------------------------------------------------
def _draws(dim, n, xi, log_u, rng, chol):
    if xi is None:
        if rng is None:
            raise ValueError("need xi or rng")
        z = rng.standard_normal((n, dim))
        xi = z @ np.asarray(chol, np.float64).T if chol is not None else z
    if log_u is None:
        if rng is None:
            raise ValueError("need log_u or rng")
        log_u = np.log(rng.random(n))
    xi = np.asarray(xi, np.float64)
    log_u = np.asarray(log_u, np.float64)
    if xi.shape[0] < n or log_u.shape[0] < n:
        raise ValueError("not enough injected randomness")
    return xi, log_u


def mh_sample_adapt(logd, x0, N, Nb=0, scale=None, xi=None, log_u=None, rng=None, chol=None,
                    return_trace=False):
    """``MH(target, proposal, x0).sample_adapt(N, Nb)`` (identification.ipynb:301-302).

    Loop of ``Ns-1`` single updates ``x* = x + scale*xi``; every ``Na=int(0.1*N)``
    steps ``lambda <- exp(log(lambda) + (mean(acc[idx:idx+Na]) - 0.234)/sqrt(i+1))``
    and ``scale = min(lambda, 1)``; ``acc[0] = 1`` belongs to the first window.

    Returns ``(samples[dim,N], target_eval[N], acc_rate, final_scale)``; with
    ``return_trace`` also the full decision vector and the proposal log-densities.
    """
    x0 = np.asarray(x0, np.float64)
    dim = x0.size
    if scale is None:
        scale = 0.1
    Ns = N + Nb
    xi, log_u = _draws(dim, Ns - 1, xi, log_u, rng, chol)
    samples = np.empty((dim, Ns))
    target_eval = np.empty(Ns)
    acc = np.zeros(Ns)
    prop_eval = np.empty(max(Ns - 1, 0))
    samples[:, 0] = x0
    target_eval[0] = logd(x0)
    acc[0] = 1
    Na = int(0.1 * N)
    lambd = scale
    star_acc = 0.234
    i = idx = 0
    for s in range(Ns - 1):
        x_star = samples[:, s] + scale * xi[s]
        t_star = logd(x_star)
        prop_eval[s] = t_star
        if decide(log_u[s], t_star, target_eval[s]):
            samples[:, s + 1], target_eval[s + 1], acc[s + 1] = x_star, t_star, 1
        else:
            samples[:, s + 1], target_eval[s + 1], acc[s + 1] = samples[:, s], target_eval[s], 0
        if Na > 0 and (s + 1) % Na == 0:
            hat = np.mean(acc[idx:idx + Na])
            zeta = 1.0 / math.sqrt(i + 1)
            lambd = math.exp(math.log(lambd) + zeta * (hat - star_acc))
            scale = min(lambd, 1.0)
            i += 1
            idx += Na
    out = (samples[:, Nb:], target_eval[Nb:], acc[Nb:].mean(), scale)
    if return_trace:
        return out + (acc, prop_eval)
    return out


def mh_sample(logd, x0, N, Nb=0, scale=1.0, xi=None, log_u=None, rng=None, chol=None,
              return_trace=False):
    """``MH.sample(N, Nb)``: the same loop without the adaptation block."""
    x0 = np.asarray(x0, np.float64)
    dim = x0.size
    Ns = N + Nb
    xi, log_u = _draws(dim, Ns - 1, xi, log_u, rng, chol)
    samples = np.empty((dim, Ns))
    target_eval = np.empty(Ns)
    acc = np.zeros(Ns)
    prop_eval = np.empty(max(Ns - 1, 0))
    samples[:, 0] = x0
    target_eval[0] = logd(x0)
    acc[0] = 1
    for s in range(Ns - 1):
        x_star = samples[:, s] + scale * xi[s]
        t_star = logd(x_star)
        prop_eval[s] = t_star
        if decide(log_u[s], t_star, target_eval[s]):
            samples[:, s + 1], target_eval[s + 1], acc[s + 1] = x_star, t_star, 1
        else:
            samples[:, s + 1], target_eval[s + 1], acc[s + 1] = samples[:, s], target_eval[s], 0
    out = (samples[:, Nb:], target_eval[Nb:], acc[Nb:].mean(), scale)
    if return_trace:
        return out + (acc, prop_eval)
    return out


# --------------------------------------------------------------------------
# component-wise MH (cuqi.sampler.CWMH semantics; never used by the reference:
# PARITY UNPINNED)
# --------------------------------------------------------------------------
def cwmh_sample(logd, x0, N, Nb=0, scale=1.0, z=None, log_u=None, rng=None, adapt=False,
                return_trace=False):
    """One step: draw ``x_o = x + scale (.) z`` once, then for j = 0..d-1 replace
    component j, evaluate, accept/reject with its own uniform; ``l`` updates
    immediately.  ``adapt``: per-component vanishing adaptation with
    ``Na=int(0.012*N)``, target ``0.21/d + 0.23``.

    z: ``[Ns-1, d]`` standard normals, log_u: ``[Ns-1, d]``.
    """
    x0 = np.asarray(x0, np.float64)
    d = x0.size
    Ns = N + Nb
    if z is None:
        z = rng.standard_normal((Ns - 1, d))
    if log_u is None:
        log_u = np.log(rng.random((Ns - 1, d)))
    z = np.asarray(z, np.float64)
    log_u = np.asarray(log_u, np.float64)
    scale = np.broadcast_to(np.asarray(scale, np.float64), (d,)).copy()
    samples = np.empty((d, Ns))
    target_eval = np.empty(Ns)
    acc = np.zeros((d, Ns))
    prop_eval = np.empty((max(Ns - 1, 0), d))
    samples[:, 0] = x0
    target_eval[0] = logd(x0)
    acc[:, 0] = 1
    Na = int(0.012 * N) if adapt else 0
    lambd = scale.copy()
    star_acc = 0.21 / d + 0.23
    i = idx = 0
    for s in range(Ns - 1):
        x_t = samples[:, s].copy()
        t_t = target_eval[s]
        x_o = x_t + scale * z[s]
        x_star = x_t.copy()
        for j in range(d):
            x_star[j] = x_o[j]
            t_star = logd(x_star)
            prop_eval[s, j] = t_star
            if decide(log_u[s, j], t_star, t_t):
                x_t[j] = x_o[j]
                t_t = t_star
                acc[j, s + 1] = 1
            x_star = x_t.copy()
        samples[:, s + 1] = x_t
        target_eval[s + 1] = t_t
        if Na > 0 and (s + 1) % Na == 0:
            hat = np.mean(acc[:, idx:idx + Na], axis=1)
            zeta = 1.0 / math.sqrt(i + 1)
            lambd = np.exp(np.log(lambd) + zeta * (hat - star_acc))
            scale = np.minimum(lambd, 1.0)
            i += 1
            idx += Na
    out = (samples[:, Nb:], target_eval[Nb:], acc[:, Nb:].mean(axis=1), scale)
    if return_trace:
        return out + (acc, prop_eval)
    return out


# --------------------------------------------------------------------------
# delayed acceptance (Christen & Fox 2005; not in CUQIpy 0.8.0 nor the
# reference: PARITY UNPINNED)
# --------------------------------------------------------------------------
def da_sample(logd_fine, logd_coarse, x0, N, Nb=0, scale=1.0, xi=None, log_u=None, rng=None,
              chol=None, return_trace=False, adapt=False):
    """Two-stage MH.  Stage 1: promote iff ``log_u1 <= min(0, lc* - lc)``.
    Stage 2 (promoted only): accept iff ``log_u2 <= min(0, (lf* - lf) - (lc* - lc))``.
    ``adapt``: the vanishing adaptation of ``MH._sample_adapt`` (window ``int(0.1*N)``, target 0.234) on the
    final accept decisions -- the definition the device library uses for adaptive delayed acceptance.

    log_u: ``[Ns-1, 2]``.  Returns samples, fine target, acceptance rate, promote rate (``adapt``: final scale appended).
    """
    x0 = np.asarray(x0, np.float64)
    dim = x0.size
    Ns = N + Nb
    if xi is None:
        zz = rng.standard_normal((Ns - 1, dim))
        xi = zz @ np.asarray(chol, np.float64).T if chol is not None else zz
    if log_u is None:
        log_u = np.log(rng.random((Ns - 1, 2)))
    xi = np.asarray(xi, np.float64)
    log_u = np.asarray(log_u, np.float64)
    samples = np.empty((dim, Ns))
    tf = np.empty(Ns)
    tc = np.empty(Ns)
    acc = np.zeros(Ns)
    prom = np.zeros(Ns)
    samples[:, 0] = x0
    tf[0] = logd_fine(x0)
    tc[0] = logd_coarse(x0)
    acc[0] = 1
    prom[0] = 1
    Na = int(0.1 * N) if adapt else 0
    lambd = scale
    i = idx = 0
    for s in range(Ns - 1):
        x_star = samples[:, s] + scale * xi[s]
        c_star = logd_coarse(x_star)
        samples[:, s + 1], tf[s + 1], tc[s + 1] = samples[:, s], tf[s], tc[s]
        if decide(log_u[s, 0], c_star, tc[s]):
            prom[s + 1] = 1
            f_star = logd_fine(x_star)
            # second-stage ratio: [pi(x*) pi~(x)] / [pi(x) pi~(x*)]
            if decide(log_u[s, 1], (f_star - tf[s]) - (c_star - tc[s]), 0.0):
                samples[:, s + 1], tf[s + 1], tc[s + 1], acc[s + 1] = x_star, f_star, c_star, 1
        if Na > 0 and (s + 1) % Na == 0:
            hat = np.mean(acc[idx:idx + Na])
            lambd = math.exp(math.log(lambd) + (hat - 0.234) / math.sqrt(i + 1))
            scale = min(lambd, 1.0)
            i += 1
            idx += Na
    out = (samples[:, Nb:], tf[Nb:], acc[Nb:].mean(), prom[Nb:].mean())
    if adapt:
        out = out + (scale,)
    if return_trace:
        return out + (acc, prom)
    return out
