"""Multi-GPU plumbing: chains shard across ranks with no communication during the run
(SURVEY.md §8e); collectives appear only at the end -- one all-reduce of split-R-hat sufficient
statistics and one all-gather of the thinned samples.  ``torch.distributed`` (NCCL on GPUs over
NVLink/NVSwitch, gloo in CPU tests) is the transport; nothing here is on the per-update path.

Replaces, at scale, the notebook loop over 5 start points and ``Samples.compute_rhat``
(tests/Xaccelerometer_geometric/identification.ipynb:298-311).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_chains_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of global chain ids owned by ``rank`` (Philox counters use global ids,
    so the union of all shards equals a single-GPU run)."""
    base, rem = divmod(n_chains_total, world)
    c0 = rank * base + min(rank, rem)
    return c0, c0 + base + (1 if rank < rem else 0)


def half_chain_moments(samples: torch.Tensor) -> torch.Tensor:
    """samples [n_keep, P, C] -> sufficient statistics [P, 4] of this shard's 2*C half-chains:
    (count, sum of half means, sum of squared half means, sum of unbiased half variances)."""
    n = samples.shape[0] // 2
    if n < 2:
        raise ValueError("need at least 4 kept draws per chain")
    P = samples.shape[1]
    if samples.is_cuda:                                              # where the samples live: one CUDA kernel
        import ctypes as C
        from . import _lib
        smp = samples.contiguous().double()
        out = torch.empty(4, P, smp.shape[2], dtype=torch.float64, device=smp.device)
        _lib.check(_lib.lib().mems_chain_moments(smp.device.index or 0, C.c_void_p(smp.data_ptr()), smp.shape[0], P,
                                                 smp.shape[2], C.c_void_p(out.data_ptr()),
                                                 C.c_void_p(torch.cuda.current_stream(smp.device).cuda_stream)), "mems_chain_moments")
        mean, var = out[0::2], out[1::2]                             # [2, P, C] each
    else:                                                            # host tensors (gloo tests of the reduction logic)
        halves = torch.stack((samples[:n], samples[-n:]))            # [2, n, P, C]
        mean = halves.mean(dim=1)
        var = halves.var(dim=1, unbiased=True)
    cnt = torch.full((P,), float(2 * samples.shape[2]), dtype=torch.float64, device=samples.device)
    return torch.stack((cnt, mean.sum(dim=(0, 2)), (mean * mean).sum(dim=(0, 2)), var.sum(dim=(0, 2))), dim=1).double()


def split_rhat(samples: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Classic split-R-hat per parameter over every chain of every rank: one all-reduce of 4*P doubles."""
    stats = half_chain_moments(samples)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    n = samples.shape[0] // 2
    m, s1, s2, sv = stats[:, 0], stats[:, 1], stats[:, 2], stats[:, 3]
    between = n * (s2 - s1 * s1 / m) / (m - 1.0)
    within = sv / m
    return torch.sqrt((between / within + n - 1.0) / n)


def posterior_moments(samples: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
    """Pooled posterior mean / std per parameter across all ranks (one all-reduce of 3*P doubles)."""
    x = samples.double()
    P = x.shape[1]
    cnt = torch.full((P,), float(x.shape[0] * x.shape[2]), dtype=torch.float64, device=x.device)
    stats = torch.stack((cnt, x.sum(dim=(0, 2)), (x * x).sum(dim=(0, 2))), dim=1)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    mean = stats[:, 1] / stats[:, 0]
    var = stats[:, 2] / stats[:, 0] - mean * mean
    return mean, torch.sqrt(torch.clamp(var, min=0.0))


def gather_samples(samples: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather the thinned samples along the chain axis: [n_keep, P, C_local] -> [n_keep, P, C_total].
    Shards may be ragged (C_total not divisible by the world size)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return samples
    world = dist.get_world_size(group)
    local = torch.tensor([samples.shape[2]], dtype=torch.int64, device=samples.device)
    sizes = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(sizes, local, group=group)
    sizes = [int(s.item()) for s in sizes]
    cmax = max(sizes)
    buf = samples.permute(2, 0, 1).contiguous()                      # chain-major for the exchange
    if buf.shape[0] < cmax:
        pad = torch.zeros((cmax - buf.shape[0],) + tuple(buf.shape[1:]), dtype=buf.dtype, device=buf.device)
        buf = torch.cat((buf, pad))
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)]).permute(1, 2, 0).contiguous()


def ess(samples: torch.Tensor, max_lag: int = 0, group: Optional[dist.ProcessGroup] = None):
    """Multi-chain effective sample size per parameter over every chain of every rank.

    The lag-autocovariance sums are computed on the device that holds the shard (``mems_chain_autocov``), one
    all-reduce of ``P*(max_lag+2)+1`` doubles combines the ranks, Geyer's truncation runs on the device too
    (``mems_ess_geyer``; host tensors -- the gloo tests -- take the numpy route).  Same estimator as arviz's ``ess`` without rank-normalisation (``diagnostics.ess_classic``).
    Returns ``(ess [P] numpy, truncated [P] bool)``."""
    import numpy as np
    from . import diagnostics
    n, P, C_local = samples.shape
    L = int(max_lag) if max_lag else min(n, 1024)
    if samples.is_cuda:
        from .engine import chain_autocov
        sums = chain_autocov(samples, L)
    else:                                                            # host tensors: gloo tests of the reduction logic
        x = samples.double()
        c = x - x.mean(dim=0, keepdim=True)
        ac = torch.stack([(c[:n - t] * c[t:]).sum(dim=0).sum(dim=1) / n for t in range(L)], dim=1)      # [P, L]
        m = x.mean(dim=0)                                            # [P, C]
        sums = torch.cat((ac, m.sum(dim=1, keepdim=True), (m * m).sum(dim=1, keepdim=True)), dim=1)
    flat = torch.cat((sums.reshape(-1), torch.tensor([float(C_local)], dtype=torch.float64, device=sums.device)))
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    Ct = int(round(float(flat[-1].item())))
    if flat.is_cuda and L >= 2 and n >= 4:                           # Geyer's truncation where the sums live (mems_ess_geyer)
        import ctypes as C
        from . import _lib
        tab = flat[:-1].reshape(P, L + 2).contiguous()
        out = torch.empty(P, dtype=torch.float64, device=tab.device)
        trunc = torch.empty(P, dtype=torch.int32, device=tab.device)
        _lib.check(_lib.lib().mems_ess_geyer(tab.device.index or 0, C.c_void_p(tab.data_ptr()), P, L, n, Ct, C.c_void_p(out.data_ptr()),
                                             C.c_void_p(trunc.data_ptr()), C.c_void_p(torch.cuda.current_stream(tab.device).cuda_stream)),
                   "mems_ess_geyer")
        return out.cpu().numpy(), trunc.cpu().numpy().astype(bool)
    tab = flat[:-1].cpu().numpy().reshape(P, L + 2)                  # host tensors: gloo tests of the reduction logic
    out, trunc = np.zeros(P), np.zeros(P, bool)
    for p in range(P):
        s1, s2 = tab[p, L], tab[p, L + 1]
        var_m = (s2 - s1 * s1 / Ct) / (Ct - 1.0) if Ct > 1 else 0.0
        out[p], trunc[p] = diagnostics.ess_from_autocov(tab[p, :L] / Ct, n, Ct, var_m)
    return out, trunc
