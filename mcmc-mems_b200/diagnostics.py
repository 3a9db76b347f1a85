"""Chain diagnostics with the semantics ``cuqi.samples.Samples.compute_ess / compute_rhat``
delegate to (arviz defaults: bulk ESS; rank-normalised split R-hat, Vehtari et al. 2021).
Call sites in the reference: tests/Xaccelerometer_geometric/identification.ipynb:308,311.

Host-side numpy (runs once after sampling); the sufficient statistics for the classic split
R-hat at scale are reduced across GPUs in ``distributed.py``.
"""
from __future__ import annotations

import numpy as np
from scipy import stats


def _split_chains(a):
    """[chain, draw] -> [2*chain, draw//2]: each chain cut into its first and last half."""
    half = a.shape[1] // 2
    return np.vstack((a[:, :half], a[:, -half:]))


def _z_scale(a):
    """Rank-normalise over all draws: ranks -> (r - 3/8)/(S + 1/4) -> normal quantiles."""
    r = stats.rankdata(a, method="average", axis=None).reshape(a.shape)
    return stats.norm.ppf((r - 0.375) / (a.size + 0.25))


def _autocov(a):
    """Biased autocovariance along axis 1 by FFT, normalised by the number of draws."""
    n = a.shape[1]
    m = 1 << int(np.ceil(np.log2(2 * n)))
    c = a - a.mean(axis=1, keepdims=True)
    f = np.fft.rfft(c, n=m, axis=1)
    return np.fft.irfft(f * np.conj(f), n=m, axis=1)[:, :n] / n


def _ess(a):
    """ESS of [chain, draw] with Geyer's initial positive + monotone sequence truncation."""
    a = np.asarray(a, np.float64)
    n_chain, n = a.shape
    if n < 4:
        return float("nan")
    acov = _autocov(a)
    chain_mean = a.mean(axis=1)
    mean_var = acov[:, 0].mean() * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if n_chain > 1:
        var_plus += chain_mean.var(ddof=1)
    rho = np.zeros(n)
    even = 1.0
    rho[0] = even
    odd = 1.0 - (mean_var - acov[:, 1].mean()) / var_plus
    rho[1] = odd
    t = 1
    while t < n - 3 and even + odd > 0.0:
        even = 1.0 - (mean_var - acov[:, t + 1].mean()) / var_plus
        odd = 1.0 - (mean_var - acov[:, t + 2].mean()) / var_plus
        if even + odd >= 0:
            rho[t + 1], rho[t + 2] = even, odd
        t += 2
    max_t = t - 2
    if even > 0:
        rho[max_t + 1] = even
    t = 1
    while t <= max_t - 2:
        if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    total = n_chain * n
    tau = -1.0 + 2.0 * rho[:max_t + 1].sum() + rho[max_t + 1:max_t + 2].sum()
    tau = max(tau, 1.0 / np.log10(total))
    return total / tau


def ess_classic(chains):
    """Multi-chain ESS of [chain, draw] without splitting or rank-normalisation (the at-scale form)."""
    return _ess(np.atleast_2d(np.asarray(chains, np.float64)))


def ess_from_autocov(acov_mean, n_draws: int, n_chains: int, chain_mean_var: float):
    """The same estimator as :func:`_ess` from its sufficient statistics: ``acov_mean[t]`` = mean over chains of
    the biased lag-t autocovariance (t < L), ``chain_mean_var`` = unbiased variance of the chain means.
    Returns ``(ess, truncated)``; ``truncated`` is True when Geyer's sequence had not turned negative by lag L."""
    acov = np.asarray(acov_mean, np.float64)
    n, L = int(n_draws), len(acov)
    mean_var = acov[0] * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if n_chains > 1:
        var_plus += chain_mean_var
    rho = np.zeros(L + 2)
    even = 1.0
    rho[0] = even
    odd = 1.0 - (mean_var - acov[1]) / var_plus
    rho[1] = odd
    t = 1
    limit = min(n - 3, L - 2)
    while t < limit and even + odd > 0.0:
        even = 1.0 - (mean_var - acov[t + 1]) / var_plus
        odd = 1.0 - (mean_var - acov[t + 2]) / var_plus
        if even + odd >= 0:
            rho[t + 1], rho[t + 2] = even, odd
        t += 2
    truncated = bool(even + odd > 0.0 and limit < n - 3)
    max_t = t - 2
    if even > 0:
        rho[max_t + 1] = even
    t = 1
    while t <= max_t - 2:
        if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    total = n_chains * n
    tau = -1.0 + 2.0 * rho[:max_t + 1].sum() + rho[max_t + 1:max_t + 2].sum()
    tau = max(tau, 1.0 / np.log10(total))
    return total / tau, truncated


def ess_bulk(chains):
    """chains: [chain, draw] of one scalar parameter -> bulk ESS (arviz ``method='bulk'``)."""
    return _ess(_z_scale(_split_chains(np.atleast_2d(np.asarray(chains, np.float64)))))


def _rhat(a):
    n = a.shape[1]
    between = n * a.mean(axis=1).var(ddof=1)
    within = a.var(axis=1, ddof=1).mean()
    return float(np.sqrt((between / within + n - 1.0) / n))


def rhat_rank(chains):
    """Rank-normalised split R-hat: max of the bulk and the folded (tail) statistic."""
    s = _split_chains(np.atleast_2d(np.asarray(chains, np.float64)))
    bulk = _rhat(_z_scale(s))
    tail = _rhat(_z_scale(np.abs(s - np.median(s))))
    return max(bulk, tail)


def split_rhat_from_moments(half_means, half_vars, n_half):
    """Classic split R-hat from per-half-chain means / unbiased variances ([m] each) -- the
    form used at scale, where only these sufficient statistics cross GPUs."""
    half_means = np.asarray(half_means, np.float64)
    half_vars = np.asarray(half_vars, np.float64)
    between = n_half * half_means.var(ddof=1)
    within = half_vars.mean()
    return float(np.sqrt((between / within + n_half - 1.0) / n_half))
