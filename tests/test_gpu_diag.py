"""Device diagnostics kernels (B200): sort, tie-averaged rank -> z, Geyer truncation, and the rank-normalised split
R-hat / bulk ESS built from them, single GPU and -- when the box has two -- across two NCCL ranks.

Reference: ``Samples.compute_ess`` / ``compute_rhat`` (tests/Xaccelerometer_geometric/identification.ipynb:308,311), i.e.
arviz's defaults.  arviz is not installable here: the host estimators in ``mcmc_mems_b200.diagnostics`` restate it from the
published algorithm (Vehtari et al. 2021) -- parity is against that restatement (K5: sanity only against the notebook).
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _mh_like_draws(n, P, Cn, seed):
    rng = np.random.default_rng(seed)
    x = np.zeros((n, P, Cn))
    cur = rng.standard_normal((P, Cn))
    for k in range(n):                                    # random-walk MH on N(0,1): ~45 % rejections -> exact ties
        prop = cur + 1.5 * rng.standard_normal((P, Cn))
        acc = np.log(rng.random((P, Cn))) < 0.5 * (cur ** 2 - prop ** 2)
        cur = np.where(acc, prop, cur)
        x[k] = cur
    x[:, 2, :] += np.linspace(0.0, 1.0, Cn)               # chains disagree on parameter 2
    return x


@pytest.mark.parametrize("n", [0, 1, 2, 3, 2047, 2048, 2049, 4096, 5000, 70001, 1 << 20])
def test_sort_is_exact(n):
    """mems_sort_f64 against numpy.sort, bit for bit: empty, tiny, around the 2048-element tile, not a power of two,
    large; values with many ties, negative zeros' neighbours, huge and tiny magnitudes."""
    from mcmc_mems_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(n + 1)
    x = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, size=n)
    if n > 10:
        x[rng.integers(0, n, size=n // 3)] = x[rng.integers(0, n, size=n // 3)]       # ties
        x[:3] = [1e300, -1e300, 0.0]
    cap = int(L.mems_sort_capacity(n))
    assert cap >= n and (n == 0 or cap & (cap - 1) == 0)
    buf = torch.zeros(max(cap, 1), dtype=torch.float64, device="cuda")
    buf[:n] = torch.as_tensor(x, device="cuda")
    _lib.check(L.mems_sort_f64(0, C.c_void_p(buf.data_ptr()), n, _stream()), "mems_sort_f64")
    assert np.array_equal(buf[:n].cpu().numpy(), np.sort(x))


def test_rank_normalize_matches_scipy():
    """Tie-averaged ranks + Blom transform + normal quantile: scipy.stats.rankdata / norm.ppf to 1e-12; the query need
    not be the whole pool (a shard of it, as on several GPUs)."""
    from scipy import stats
    from mcmc_mems_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(3)
    M = 30011
    pool = np.round(rng.standard_normal(M), 2)             # heavy ties
    srt = torch.as_tensor(np.sort(pool), device="cuda")
    for sl in (slice(0, M), slice(1000, 1777), slice(5, 6)):
        x = torch.as_tensor(pool[sl].copy(), device="cuda")
        z = torch.empty_like(x)
        _lib.check(L.mems_rank_normalize(0, C.c_void_p(srt.data_ptr()), M, C.c_void_p(x.data_ptr()), x.numel(),
                                         C.c_void_p(z.data_ptr()), _stream()), "mems_rank_normalize")
        r = stats.rankdata(pool, method="average")[sl]
        want = stats.norm.ppf((r - 0.375) / (M + 0.25))
        assert np.allclose(z.cpu().numpy(), want, rtol=1e-12, atol=1e-14)


def test_geyer_truncation_matches_the_host_estimator():
    """mems_ess_geyer on autocovariance sums against diagnostics.ess_from_autocov: chains that mix fast, slowly, and a lag
    window too short for the sequence to turn negative (truncated flag)."""
    from mcmc_mems_b200 import _lib, diagnostics as D
    L = _lib.lib()
    rng = np.random.default_rng(9)
    n, m = 500, 7
    cases = []
    for phi in (0.0, 0.6, 0.97):
        e = rng.standard_normal((m, n))
        x = np.zeros((m, n))
        for k in range(1, n):
            x[:, k] = phi * x[:, k - 1] + e[:, k]
        cases.append(x)
    for Lag in (n, 40, 6):
        tab = np.zeros((len(cases), Lag + 2))
        for i, x in enumerate(cases):
            tab[i, :Lag] = D._autocov(x)[:, :Lag].sum(axis=0)
            tab[i, Lag] = x.mean(axis=1).sum()
            tab[i, Lag + 1] = (x.mean(axis=1) ** 2).sum()
        t = torch.as_tensor(tab, device="cuda")
        out = torch.empty(len(cases), dtype=torch.float64, device="cuda")
        tr = torch.empty(len(cases), dtype=torch.int32, device="cuda")
        _lib.check(L.mems_ess_geyer(0, C.c_void_p(t.data_ptr()), len(cases), Lag, n, m, C.c_void_p(out.data_ptr()),
                                    C.c_void_p(tr.data_ptr()), _stream()), "mems_ess_geyer")
        for i, x in enumerate(cases):
            s1, s2 = tab[i, Lag], tab[i, Lag + 1]
            want, wtr = D.ess_from_autocov(tab[i, :Lag] / m, n, m, (s2 - s1 * s1 / m) / (m - 1.0))
            assert out[i].item() == pytest.approx(want, rel=1e-12)
            assert bool(tr[i].item()) == wtr
            if Lag == n:
                assert want == pytest.approx(D.ess_classic(x), rel=1e-9)


@pytest.mark.parametrize("n", [400, 401, 9])
def test_rank_normalised_rhat_and_ess_match_the_host_estimators(n):
    """The device pipeline (sort -> rank -> z -> moments / lag sums -> Geyer) against the arviz-semantics host estimators,
    even and odd draw counts, on chains full of ties."""
    from mcmc_mems_b200 import diagnostics as D, diagnostics_device as DD
    x = _mh_like_draws(n, 3, 23, 17)
    smp = torch.as_tensor(x, device="cuda")
    ess, rhat = DD.ess_bulk(smp, max_lag=n // 2), DD.rhat_rank(smp)
    for p in range(3):
        chains = x[:, p, :].T
        assert ess[p] == pytest.approx(D.ess_bulk(chains), rel=1e-9)
        assert rhat[p] == pytest.approx(D.rhat_rank(chains), rel=1e-9)


def _nccl_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from mcmc_mems_b200 import BatchedSamples, distributed as DS, diagnostics_device as DD
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        out = {}
        for n in (120, 121):
            full = _mh_like_draws(n, 3, 37, 5)
            c0, c1 = DS.shard_range(full.shape[2], rank, world)
            mine = torch.as_tensor(full[:, :, c0:c1].copy(), device=f"cuda:{rank}")
            b = BatchedSamples(None, None, np.zeros(c1 - c0), np.ones(c1 - c0), device_samples=mine)
            out[n] = (b.compute_rhat(on_device=True, device=rank, across_ranks=True),
                      DD.ess_bulk(mine, max_lag=n // 2), b.compute_ess(device=rank, across_ranks=True))
        ret[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_rank_normalised_diagnostics_equal_the_host_estimators():
    """BatchedSamples.compute_rhat / compute_ess(across_ranks=True) on two GPUs: one all-gather of the pool, own-kernel sort
    and ranks on each GPU, one all-reduce of the moments / lag sums -- equal to the single-process host estimators to 1e-9."""
    import torch.multiprocessing as mp
    from mcmc_mems_b200 import diagnostics as D
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_nccl_worker, args=(world, port, ret), nprocs=world, join=True)
    for n in (120, 121):
        full = _mh_like_draws(n, 3, 37, 5)
        for p in range(3):
            chains = full[:, p, :].T
            for r in range(world):
                assert ret[r][n][0][p] == pytest.approx(D.rhat_rank(chains), rel=1e-9)
                assert ret[r][n][1][p] == pytest.approx(D.ess_bulk(chains), rel=1e-9)
        assert np.all(np.isfinite(ret[0][n][2]))
