"""arviz-default diagnostics (rank-normalised split R-hat, bulk ESS) for large batches, on the device(s) that hold the
samples.  Same estimators as ``diagnostics.rhat_rank`` / ``diagnostics.ess_bulk`` (what ``Samples.compute_rhat`` /
``compute_ess`` report in the reference, tests/Xaccelerometer_geometric/identification.ipynb:308,311).

Every arithmetic step is one of the library's own kernels (``csrc/mems_diag.cu``, ``mems_aux.cu``, ``mems_fp32.cu``):

    pool      the draws of every chain, per parameter; with several GPUs one all-gather makes the pool global
    sort      ``mems_sort_f64``  (bitonic network)
    rank, z   ``mems_rank_normalize``  (tie-averaged ranks by binary search in the sorted pool -- a Metropolis chain
              repeats its state at every rejection, so ties are the rule -- then the Blom transform and Phi^-1)
    R-hat     ``mems_chain_moments`` of z (half-chain means / variances), one all-reduce of 4*P doubles
    ESS       ``mems_chain_autocov`` of z (lag sums), one all-reduce, ``mems_ess_geyer`` (Geyer truncation)

Ranks are global, so the numbers do not depend on how the chains are sharded over GPUs.  Host tensors take a numpy
route with the same structure: it exists for the gloo tests of the collective logic, the product path is the device one.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, diagnostics


def _dist_on(group) -> bool:
    """``group=None``: the default process group when one is initialised; ``group=False``: this process only."""
    if group is False:
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _dev(t: torch.Tensor) -> int:
    return t.device.index or 0


def _split_draws(samples: torch.Tensor) -> torch.Tensor:
    """[n, P, C] -> [2*(n//2), P, C]: the draws arviz's ``_split_chains`` keeps (an odd count drops the middle draw);
    rows [:half] are the first half-chains, rows [half:] the last."""
    half = samples.shape[0] // 2
    if 2 * half == samples.shape[0]:
        return samples.contiguous()
    return torch.cat((samples[:half], samples[-half:]), dim=0).contiguous()


def _gather_ragged(flat: torch.Tensor, group) -> torch.Tensor:
    """[P, m_local] on every rank -> [P, M] (all ranks' columns, rank order)."""
    world = dist.get_world_size(group)
    local = torch.tensor([flat.shape[1]], dtype=torch.int64, device=flat.device)
    sizes = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(sizes, local, group=group)
    sizes = [int(s.item()) for s in sizes]
    mmax = max(sizes)
    buf = flat
    if flat.shape[1] < mmax:
        buf = torch.cat((flat, torch.zeros(flat.shape[0], mmax - flat.shape[1], dtype=flat.dtype, device=flat.device)), dim=1)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf.contiguous(), group=group)
    return torch.cat([o[:, :s] for o, s in zip(out, sizes)], dim=1)


def sorted_pool(flat: torch.Tensor, group=None) -> torch.Tensor:
    """flat [P, m] (this rank's draws per parameter) -> the pooled draws of all ranks, sorted ascending: [P, M]."""
    pool = _gather_ragged(flat, group) if _dist_on(group) else flat
    P, M = pool.shape
    if not pool.is_cuda:
        return torch.from_numpy(np.sort(pool.numpy(), axis=1))
    L = _lib.lib()
    cap = int(L.mems_sort_capacity(M))
    buf = torch.empty(P, cap, dtype=torch.float64, device=pool.device)
    buf[:, :M] = pool
    for p in range(P):
        _lib.check(L.mems_sort_f64(_dev(buf), C.c_void_p(buf[p].data_ptr()), M, _stream(buf)), "mems_sort_f64")
    return buf[:, :M]


def rank_normalize(pool_sorted: torch.Tensor, flat: torch.Tensor) -> torch.Tensor:
    """z-scores of ``flat`` [P, m] against the sorted pool [P, M]: tie-averaged ranks -> (r - 3/8)/(M + 1/4) -> Phi^-1."""
    P, M = pool_sorted.shape
    if not flat.is_cuda:
        from scipy.special import ndtri
        out = np.empty(flat.shape)
        for p in range(P):
            srt, v = pool_sorted[p].numpy(), flat[p].numpy()
            lo = np.searchsorted(srt, v, side="left")
            hi = np.searchsorted(srt, v, side="right")
            out[p] = ndtri((0.5 * (lo + 1 + hi) - 0.375) / (M + 0.25))
        return torch.from_numpy(out)
    L = _lib.lib()
    flat = flat.contiguous()
    z = torch.empty_like(flat)
    for p in range(P):
        row = pool_sorted[p]
        if not row.is_contiguous():
            row = row.contiguous()
        _lib.check(L.mems_rank_normalize(_dev(flat), C.c_void_p(row.data_ptr()), M, C.c_void_p(flat[p].data_ptr()),
                                         flat.shape[1], C.c_void_p(z[p].data_ptr()), _stream(flat)), "mems_rank_normalize")
    return z


def z_scale(x: torch.Tensor, group=None, pool_sorted: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Rank-normalise x [n, P, C] over all draws of all chains (of all ranks), per parameter -> z [n, P, C]."""
    n, P, Cn = x.shape
    flat = x.permute(1, 0, 2).reshape(P, n * Cn).contiguous()
    if pool_sorted is None:
        pool_sorted = sorted_pool(flat, group)
    return rank_normalize(pool_sorted, flat).reshape(P, n, Cn).permute(1, 0, 2).contiguous()


def _pool_median(pool_sorted: torch.Tensor) -> torch.Tensor:
    """numpy's median (mean of the two middle values of an even count) of every row of the sorted pool."""
    M = pool_sorted.shape[1]
    return 0.5 * (pool_sorted[:, (M - 1) // 2] + pool_sorted[:, M // 2])


def _rhat_of_split_z(z: torch.Tensor, group) -> torch.Tensor:
    """z [2*half, P, C] (first halves then last halves) -> classic R-hat over the 2*C half-chains of every rank."""
    from .distributed import half_chain_moments
    stats = half_chain_moments(z)                                      # [P, 4]: count, sum mean, sum mean^2, sum var
    if _dist_on(group):
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    n = z.shape[0] // 2
    m, s1, s2, sv = stats[:, 0], stats[:, 1], stats[:, 2], stats[:, 3]
    between = n * (s2 - s1 * s1 / m) / (m - 1.0)
    within = sv / m
    return torch.sqrt((between / within + n - 1.0) / n)


def rhat_rank(samples: torch.Tensor, group=None) -> np.ndarray:
    """Rank-normalised split R-hat per parameter (max of the bulk and the folded statistic) over every chain of every
    rank.  samples [n, P, C_local]."""
    s = _split_draws(samples.double())
    n2, P, Cn = s.shape
    flat = s.permute(1, 0, 2).reshape(P, n2 * Cn).contiguous()
    pool = sorted_pool(flat, group)
    bulk = _rhat_of_split_z(z_scale(s, group, pool), group)
    folded = (s - _pool_median(pool)[None, :, None]).abs()
    tail = _rhat_of_split_z(z_scale(folded, group), group)
    return torch.maximum(bulk, tail).cpu().numpy()


def _autocov_sums(z_split: torch.Tensor, L: int) -> torch.Tensor:
    """z_split [n, P, m] -> [P, L + 2]: sums over chains of the biased lag autocovariances, of the chain means and of
    their squares."""
    if z_split.is_cuda:
        from .engine import chain_autocov
        return chain_autocov(z_split, L)
    n = z_split.shape[0]
    c = z_split - z_split.mean(dim=0, keepdim=True)
    ac = torch.stack([(c[:n - t] * c[t:]).sum(dim=0).sum(dim=1) / n for t in range(L)], dim=1)
    m = z_split.mean(dim=0)
    return torch.cat((ac, m.sum(dim=1, keepdim=True), (m * m).sum(dim=1, keepdim=True)), dim=1)


def ess_bulk(samples: torch.Tensor, max_lag: int = 0, group=None, return_truncated: bool = False):
    """Bulk effective sample size per parameter (arviz ``method='bulk'``) over every chain of every rank.
    samples [n, P, C_local]."""
    s = _split_draws(samples.double())
    z = z_scale(s, group)
    half = z.shape[0] // 2
    P = z.shape[1]
    z_split = torch.cat((z[:half], z[half:]), dim=2).contiguous()       # [half, P, 2C]: every half-chain is a chain
    L = int(max_lag) if max_lag else min(half, 1024)
    sums = _autocov_sums(z_split, L)
    flat = torch.cat((sums.reshape(-1), torch.tensor([float(z_split.shape[2])], dtype=torch.float64, device=sums.device)))
    if _dist_on(group):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    m_total = int(round(float(flat[-1].item())))
    tab = flat[:-1].reshape(P, L + 2).contiguous()
    if tab.is_cuda:
        out = torch.empty(P, dtype=torch.float64, device=tab.device)
        trunc = torch.empty(P, dtype=torch.int32, device=tab.device)
        _lib.check(_lib.lib().mems_ess_geyer(_dev(tab), C.c_void_p(tab.data_ptr()), P, L, half, m_total, C.c_void_p(out.data_ptr()),
                                             C.c_void_p(trunc.data_ptr()), _stream(tab)), "mems_ess_geyer")
        ess, tr = out.cpu().numpy(), trunc.cpu().numpy().astype(bool)
    else:
        tabn = tab.numpy()
        ess, tr = np.zeros(P), np.zeros(P, bool)
        for p in range(P):
            s1, s2 = tabn[p, L], tabn[p, L + 1]
            var_m = (s2 - s1 * s1 / m_total) / (m_total - 1.0) if m_total > 1 else 0.0
            ess[p], tr[p] = diagnostics.ess_from_autocov(tabn[p, :L] / m_total, half, m_total, var_m)
    return (ess, tr) if return_truncated else ess

