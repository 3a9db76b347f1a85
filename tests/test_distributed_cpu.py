"""world_size-2 gloo tests of the multi-GPU host logic (sharding, R-hat reduction, sample gather)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from mcmc_mems_b200 import distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        C_total, n_keep, P = 11, 40, 3                       # ragged on purpose
        g = torch.Generator().manual_seed(0)
        full = torch.randn(n_keep, P, C_total, generator=g, dtype=torch.float64)
        full[:, 2, :] += torch.linspace(0, 3, C_total)       # chains disagree on parameter 2
        c0, c1 = D.shard_range(C_total, rank, world)
        mine = full[:, :, c0:c1].contiguous()
        rhat = D.split_rhat(mine)
        mean, std = D.posterior_moments(mine)
        gathered = D.gather_samples(mine)
        ess, trunc = D.ess(mine, max_lag=n_keep)
        ret[rank] = (rhat.numpy(), mean.numpy(), std.numpy(), gathered.numpy(), full.numpy(), (c0, c1), ess, trunc)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_rhat_and_gather():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    from mcmc_mems_b200 import diagnostics
    r0, r1 = ret[0], ret[1]
    full = r0[4]
    assert r0[5] == (0, 6) and r1[5] == (6, 11)
    for r in (r0, r1):
        assert np.array_equal(r[3], full)                    # gather restores global chain order
        assert np.allclose(r[1], full.transpose(1, 0, 2).reshape(3, -1).mean(axis=1))
        assert np.allclose(r[2], full.transpose(1, 0, 2).reshape(3, -1).std(axis=1))
    n = full.shape[0] // 2
    for p in range(3):
        halves = np.concatenate([full[:n, p, :].T, full[-n:, p, :].T])
        want = diagnostics.split_rhat_from_moments(halves.mean(axis=1), halves.var(axis=1, ddof=1), n)
        assert r0[0][p] == pytest.approx(want, rel=1e-12) and r1[0][p] == pytest.approx(want, rel=1e-12)
    assert r0[0][2] > 1.2 and abs(r0[0][0] - 1) < 0.1
    for p in range(3):                                       # ESS over all chains of both ranks == single-process estimator
        want = diagnostics.ess_classic(full[:, p, :].T)
        assert r0[6][p] == pytest.approx(want, rel=1e-10) and r1[6][p] == pytest.approx(want, rel=1e-10)


def _mh_like_draws(n, P, Cn, seed):
    """Random-walk MH on N(0,1): ~45 % rejections -> exact ties, as in a real Metropolis chain."""
    rng = np.random.default_rng(seed)
    x = np.zeros((n, P, Cn))
    cur = rng.standard_normal((P, Cn))
    for k in range(n):
        prop = cur + 1.5 * rng.standard_normal((P, Cn))
        acc = np.log(rng.random((P, Cn))) < 0.5 * (cur ** 2 - prop ** 2)
        cur = np.where(acc, prop, cur)
        x[k] = cur
    x[:, 2, :] += np.linspace(0.0, 1.0, Cn)               # chains disagree on parameter 2
    return x


def _rank_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from mcmc_mems_b200 import distributed as D, diagnostics_device as DD
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = {}
        for n in (60, 61):                                   # even and odd draw counts (odd: arviz drops the middle draw)
            full = torch.from_numpy(_mh_like_draws(n, 3, 11, 5))
            c0, c1 = D.shard_range(full.shape[2], rank, world)
            mine = full[:, :, c0:c1].contiguous()
            out[n] = (DD.rhat_rank(mine), DD.ess_bulk(mine, max_lag=n // 2), DD.rhat_rank(mine, group=False))
        ret[rank] = out
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_rank_normalised_diagnostics_equal_the_host_estimators():
    """Rank-normalised split R-hat and bulk ESS over shards on two ranks (global ranks through one all-gather of the
    pool, one all-reduce of the moments / lag sums) equal the single-process arviz-semantics estimators."""
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rank_worker, args=(world, port, ret), nprocs=world, join=True)
    from mcmc_mems_b200 import diagnostics
    for n in (60, 61):
        full = _mh_like_draws(n, 3, 11, 5)
        for p in range(3):
            chains = full[:, p, :].T
            for r in (0, 1):
                assert ret[r][n][0][p] == pytest.approx(diagnostics.rhat_rank(chains), rel=1e-9)
                assert ret[r][n][1][p] == pytest.approx(diagnostics.ess_bulk(chains), rel=1e-9)
        # group=False restricts the estimate to the rank's own shard
        assert ret[0][n][2][2] == pytest.approx(diagnostics.rhat_rank(full[:, 2, :6].T), rel=1e-9)
        assert ret[0][n][0][2] > 1.05


def test_shard_range_covers_everything():
    from mcmc_mems_b200.distributed import shard_range
    for total in (1, 7, 4096, 1048576):
        for world in (1, 2, 4, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
