"""arviz-default diagnostics (rank-normalised split R-hat, bulk ESS) for large batches, on the device that holds the
samples.  Same estimators as ``diagnostics.rhat_rank`` / ``diagnostics.ess_bulk`` (what ``Samples.compute_rhat`` /
``compute_ess`` report in the reference, tests/Xaccelerometer_geometric/identification.ipynb:308,311).

Division of labour: the global sort and the normal quantile function are library calls (``torch.sort``,
``torch.special.ndtri`` -- plumbing, like a cuBLAS call); tie-averaged ranks are needed because a Metropolis chain
repeats its state at every rejection.  The lag-autocovariance sums of the rank-normalised draws come from the
library's own kernel (``mems_chain_autocov``); Geyer's truncation runs on the host on a vector of a few hundred numbers.
Single device: global ranks over several GPUs would need a distributed sort (``distributed.ess`` / ``split_rhat``
are the multi-GPU estimators).
"""
from __future__ import annotations

import numpy as np
import torch

from . import diagnostics
from ._lib import MemsError


def _split(samples: torch.Tensor) -> torch.Tensor:
    """[n, P, C] -> [n//2, P, 2C]: every chain cut into its first and last half (arviz ``_split_chains``)."""
    half = samples.shape[0] // 2
    return torch.cat((samples[:half], samples[-half:]), dim=2).contiguous()


def _z_scale(x: torch.Tensor) -> torch.Tensor:
    """Rank-normalise over all draws of all chains, per parameter: x [n, P, C] -> z [n, P, C].
    ranks (ties averaged) -> (r - 3/8)/(S + 1/4) -> standard normal quantiles."""
    n, P, C = x.shape
    flat = x.permute(1, 0, 2).reshape(P, n * C)
    M = flat.shape[1]
    vals, idx = torch.sort(flat, dim=1)
    new = torch.ones_like(vals, dtype=torch.bool)
    new[:, 1:] = vals[:, 1:] != vals[:, :-1]
    gid = torch.cumsum(new.to(torch.int64), dim=1) - 1                          # run id of every sorted element
    pos = torch.arange(1, M + 1, device=x.device, dtype=torch.float64).expand(P, M)
    sums = torch.zeros(P, M, dtype=torch.float64, device=x.device).scatter_add_(1, gid, pos)
    cnts = torch.zeros(P, M, dtype=torch.float64, device=x.device).scatter_add_(1, gid, torch.ones_like(pos))
    avg = (sums / cnts.clamp_min(1.0)).gather(1, gid)                           # tie-averaged rank, sorted order
    ranks = torch.empty_like(avg).scatter_(1, idx, avg)                          # back to draw order
    z = torch.special.ndtri((ranks - 0.375) / (M + 0.25))
    return z.reshape(P, n, C).permute(1, 0, 2).contiguous()


def _rhat(z: torch.Tensor) -> torch.Tensor:
    """Classic R-hat of z [n, P, m chains] per parameter."""
    n = z.shape[0]
    between = n * z.mean(dim=0).var(dim=1, unbiased=True)
    within = z.var(dim=0, unbiased=True).mean(dim=1)
    return torch.sqrt((between / within + n - 1.0) / n)


def rhat_rank(samples: torch.Tensor) -> np.ndarray:
    """Rank-normalised split R-hat per parameter (max of the bulk and the folded statistic).  samples [n, P, C] on a GPU."""
    if not samples.is_cuda:
        raise MemsError("diagnostics_device needs device samples")
    s = _split(samples.double())
    bulk = _rhat(_z_scale(s))
    n, P, m = s.shape
    med = s.permute(1, 0, 2).reshape(P, -1).median(dim=1).values          # lower median, as torch defines it
    # numpy's median averages the two middle values of an even count: reproduce it
    flat = s.permute(1, 0, 2).reshape(P, -1)
    k = flat.shape[1]
    if k % 2 == 0:
        srt = torch.sort(flat, dim=1).values
        med = 0.5 * (srt[:, k // 2 - 1] + srt[:, k // 2])
    tail = _rhat(_z_scale((s - med[None, :, None]).abs()))
    return torch.maximum(bulk, tail).cpu().numpy()


def ess_bulk(samples: torch.Tensor, max_lag: int = 0) -> np.ndarray:
    """Bulk effective sample size per parameter (arviz ``method='bulk'``).  samples [n, P, C] on a GPU."""
    if not samples.is_cuda:
        raise MemsError("diagnostics_device needs device samples")
    from .engine import chain_autocov
    z = _z_scale(_split(samples.double()))
    n, P, m = z.shape
    L = int(max_lag) if max_lag else min(n, 1024)
    tab = chain_autocov(z, L).cpu().numpy()
    out = np.zeros(P)
    for p in range(P):
        s1, s2 = tab[p, L], tab[p, L + 1]
        var_m = (s2 - s1 * s1 / m) / (m - 1.0) if m > 1 else 0.0
        out[p], _ = diagnostics.ess_from_autocov(tab[p, :L] / m, n, m, var_m)
    return out
