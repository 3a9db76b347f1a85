"""CPU tests: C-ABI surface, host-side logic of the drop-in API, diagnostics.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers


def test_abi_library_exports_every_declared_symbol(lib_built):
    hdr = open(os.path.join(os.path.dirname(helpers.GOLDEN), "..", "include", "mems_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mems_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 14
    from mcmc_mems_b200 import LIB_PATH, _lib
    L = ctypes.CDLL(LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/mems_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert _lib.lib().mems_version() == 100


def test_abi_argument_errors_without_gpu(lib_built):
    from mcmc_mems_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    bad = _lib.MemsConfig(9, 150, 4, 0, 0, 0)
    assert L.mems_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    assert b"n_params" in L.mems_last_error()
    bad = _lib.MemsConfig(3, 150, 4, 0, 0, 129)
    assert L.mems_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    assert L.mems_run(None, 1, 0, 1, None, None, None) == -1
    # device-diagnostics entry points: argument checks and the host-only capacity rule (power of two, >= one 2048-key tile)
    assert [L.mems_sort_capacity(n) for n in (-1, 0, 1, 2048, 2049, 70001)] == [0, 2048, 2048, 2048, 4096, 131072]
    assert L.mems_sort_f64(0, None, -1, None) == -1 and L.mems_sort_f64(0, None, 1, None) == 0      # n <= 1: nothing to do
    assert L.mems_rank_normalize(0, None, 0, None, 5, None, None) == -1
    assert L.mems_ess_geyer(0, None, 1, 8, 100, 4, None, None, None) == -1
    import torch
    if not torch.cuda.is_available():
        assert L.mems_sort_f64(0, ctypes.c_void_p(16), 5, None) == -4 and b"no CPU fallback" in L.mems_last_error()
        ok = _lib.MemsConfig(3, 150, 4, 0, 0, 0)
        assert L.mems_create(ctypes.byref(ok), ctypes.byref(h)) == -4      # no device, no CPU fallback
        assert b"no CPU fallback" in L.mems_last_error()


def test_engine_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mcmc_mems_b200 import MemsError, PATH_TC
    pb = helpers.load_problem("geometric")
    with pytest.raises(MemsError):
        helpers.engine(pb, 8, PATH_TC)
    from mcmc_mems_b200 import NN_Model
    m = NN_Model()
    m.set_layers(pb["layers"])
    with pytest.raises(MemsError):
        m.predict(np.zeros((2, 4)))


def test_spec_validation_and_cholesky():
    from mcmc_mems_b200 import MHEngine, SurrogateSpec
    pb = helpers.load_problem("geometric")
    spec = helpers.spec_of(pb)
    spec.validate()
    assert spec.n_params == 3 and spec.n_time == 150
    bad = SurrogateSpec(pb["layers"][:-1], pb["time"], pb["scaler_mean"], pb["scaler_scale"])
    with pytest.raises(ValueError):
        bad.validate()
    cov = np.zeros((4, 4))
    cov[3, 3] = 0.04                      # the "1-param Q" configuration: only Q moves
    L = MHEngine._chol_psd(cov)
    assert L[3, 3] == pytest.approx(0.2) and np.count_nonzero(L) == 1
    L = MHEngine._chol_psd(pb["cov"])
    assert np.allclose(L @ L.T, pb["cov"])


def test_cuqi_lookalikes_build_the_notebook_posterior():
    from mcmc_mems_b200 import (Continuous1D, Discrete, FixtureProcessor, Gaussian, JointDistribution, MH, Model,
                                NN_Model, Uniform)
    pb = helpers.load_problem("geometric")
    T = len(pb["time"])
    fwd = helpers.oracle_forward(pb)                  # any callable works for the host-side objects
    fwd.spec = helpers.spec_of(pb)
    model = Model(forward=fwd, range_geometry=Continuous1D(T), domain_geometry=Discrete(["Overetch", "Offset", "Thickness"]))
    x = Uniform(low=pb["low"], high=pb["high"])
    y = Gaussian(mean=model(x), cov=pb["sigma2"] * np.eye(T))
    post = JointDistribution(x, y)(y_distribution=pb["y_obs"])
    assert post.dim == 3 and post.sigma2 == pytest.approx(0.005)
    # host-side logd equals the oracle's posterior log-density
    assert post.logd(pb["x_true"]) == pytest.approx(helpers.oracle_logd(pb)(pb["x_true"]), rel=1e-12)
    assert post.logd(np.array([0.6, 0, 30.0])) == -np.inf
    with pytest.raises(ValueError):
        MH(target=post, proposal=Gaussian(mean=np.ones(3), cov=pb["cov"]), x0=pb["x_true"])   # not symmetric
    s = MH(target=post, proposal=Gaussian(mean=np.zeros(3), cov=pb["cov"]), x0=pb["x_true"])
    with pytest.raises(ValueError):
        s.sample(10)                                  # scale must be set for the non-adaptive sampler
    g = Gaussian(mean=pb["y_true"], cov=pb["sigma2"] * np.eye(T))
    d = g.sample(rng=np.random.default_rng(0))
    assert d.shape == (T,) and abs(np.std(d - pb["y_true"]) / np.sqrt(pb["sigma2"]) - 1) < 0.2
    with pytest.raises(ValueError):
        Gaussian(mean=model(x), cov=np.diag(np.linspace(1, 2, T))).iid_variance(T)
    assert isinstance(NN_Model(), NN_Model) and FixtureProcessor(pb["time"], pb["scaler_mean"], pb["scaler_scale"]).time.shape == (T,)


def test_samples_container_semantics():
    from mcmc_mems_b200 import BatchedSamples, Samples
    rng = np.random.default_rng(0)
    s = Samples(rng.standard_normal((3, 6000)), loglike_eval=np.zeros(6000), acc_rate=0.27)
    b = s.burnthin(1000, 5)
    assert b.samples.shape == (3, 1000) and np.array_equal(b.samples[:, 1], s.samples[:, 1005])
    assert s.mean().shape == (3,)
    ess = b.compute_ess()
    assert np.all(ess > 600) and np.all(ess < 1500)              # iid draws: ESS ~ N
    others = [Samples(rng.standard_normal((3, 6000))) for _ in range(4)]
    rh = s.compute_rhat(others)
    assert np.all(np.abs(rh - 1) < 0.01)
    with pytest.raises(TypeError):
        s.compute_rhat([Samples(rng.standard_normal((3, 10)))])
    B = BatchedSamples(rng.standard_normal((5, 3, 400)), None, np.full(5, 0.3), np.full(5, 0.1))
    assert len(B) == 5 and B[2].samples.shape == (3, 400) and B.burnthin(100, 2).samples.shape == (5, 3, 150)
    assert np.all(np.abs(B.compute_rhat() - 1) < 0.05)


def test_diagnostics_on_ar1():
    from mcmc_mems_b200 import diagnostics
    rng = np.random.default_rng(1)
    rho, n = 0.9, 20000
    x = np.zeros((4, n))
    e = rng.standard_normal((4, n))
    for t in range(1, n):
        x[:, t] = rho * x[:, t - 1] + e[:, t]
    ess = diagnostics.ess_bulk(x)
    want = 4 * n * (1 - rho) / (1 + rho)
    assert 0.7 * want < ess < 1.4 * want
    assert abs(diagnostics.rhat_rank(x) - 1) < 0.01
    y = x.copy()
    y[0] += 3.0
    assert diagnostics.rhat_rank(y) > 1.1
    hm = np.array([c[:n // 2].mean() for c in x] + [c[n // 2:].mean() for c in x])
    hv = np.array([c[:n // 2].var(ddof=1) for c in x] + [c[n // 2:].var(ddof=1) for c in x])
    assert abs(diagnostics.split_rhat_from_moments(hm, hv, n // 2) - 1) < 0.01


def test_ess_from_autocov_equals_the_full_estimator():
    """The at-scale ESS (sufficient statistics -> Geyer on the host) is the same estimator as the full-sample one."""
    from mcmc_mems_b200 import diagnostics as D
    rng = np.random.default_rng(3)
    n_chain, n = 6, 800
    x = np.zeros((n_chain, n))
    for k in range(1, n):
        x[:, k] = 0.8 * x[:, k - 1] + rng.standard_normal(n_chain)
    want = D.ess_classic(x)
    acov = D._autocov(x).mean(axis=0)
    got, trunc = D.ess_from_autocov(acov, n, n_chain, x.mean(axis=1).var(ddof=1))
    assert not trunc and got == pytest.approx(want, rel=1e-12)
    got2, trunc2 = D.ess_from_autocov(acov[:200], n, n_chain, x.mean(axis=1).var(ddof=1))     # lags cut at 200: enough for rho=0.8
    assert not trunc2 and got2 == pytest.approx(want, rel=1e-12)
    _, trunc3 = D.ess_from_autocov(acov[:4], n, n_chain, x.mean(axis=1).var(ddof=1))          # too few lags: flagged
    assert trunc3
    assert 0.05 * n_chain * n < want < 0.25 * n_chain * n                                   # (1-rho)/(1+rho) = 0.11


def test_sensitivity_fixture_and_push_forward_math():
    """Pin the sensitivity surrogate (config_II) the push-forward uses: the numpy restatement of the Keras stack
    reproduces the held-out targets of the reference's own split to surrogate accuracy."""
    import helpers
    from oracle.surrogate import mlp_forward
    d = np.load(os.path.join(helpers.GOLDEN, "sensitivity.npz"))
    layers = helpers.layers_of(d)
    assert [k.shape for k, _, _ in layers] == [(3, 32), (32, 32), (32, 32), (32, 32), (32, 1)]
    rows = (d["X_test"] - d["scaler_mean"]) / d["scaler_scale"]
    assert np.allclose(rows, d["X_test_scaled"], rtol=0, atol=1e-12)
    pred = mlp_forward(layers, rows.astype(np.float32)).reshape(-1)
    rel = np.abs(pred - d["y_test"]) / np.abs(d["y_test"])
    assert np.median(rel) < 0.01 and rel.max() < 0.05


def test_bench_reference_arm_pieces_run_on_cpu():
    """bench.py --impl reference times this function on every host core: keep it importable and sane without a GPU."""
    import bench
    dt = bench._cpu_chain((150, 40, 0))
    assert 0.0 < dt < 30.0
    assert bench.flop_per_row(3) == 41600 and bench.flop_per_row(4) == 41728
    class A:   # noqa: E701
        workload, chains, inner, thin, path, time_points = "geometric_4096", 4096, 256, 64, "tc", 150
    traffic, src = bench.measured_traffic(A, 4096)
    assert traffic > 100000 and src.endswith(".json")
    assert bench.measured_traffic(A, 1024) == (None, None)
    assert bench.chains_of(A, 8) == (8 * 4096, 4096)
    A.workload, A.chains = "sweep_1M", 0
    assert bench.chains_of(A, 8) == (1 << 20, 1 << 17)
    cfg = bench.workload_config(A, 4)
    assert cfg["workload"] == "sweep_1M" and cfg["chains_total"] == 1 << 20 and cfg["chains_per_gpu"] == 1 << 18
    # the bench rebuilds the reference-shaped objects from the package's fixtures, not from the test tree
    src = open(bench.__file__).read()
    assert "helpers" not in src and "os.path.join(ROOT, \"tests\")" not in src
