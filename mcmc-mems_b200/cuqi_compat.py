"""CUQIpy-shaped objects sufficient for the reference's notebook cells, backed by the batched
GPU sampler.  CUQIpy itself (``cuqipy==0.8.0``, reference README.md:36) is not vendored and not
installable here; these look-alikes keep the call shapes of

    tests/Xaccelerometer_geometric/identification.ipynb:136-139  Model(forward=, range_geometry=, domain_geometry=)
    ...:161   Gaussian(mean=y_true, cov=noise*np.eye(T)).sample()
    ...:165   Uniform(low=, high=)
    ...:169   Gaussian(mean=cuqi_model(x_distribution), cov=noise*np.eye(T))
    ...:292   JointDistribution(x_distribution, y_distribution)(y_distribution=y_observed)
    ...:300-302  MH(target=posterior, proposal=Gaussian(mean=0, cov), x0=x0).sample_adapt(N, Nb)
    ...:308,311  Samples.burnthin / compute_ess / compute_rhat

Extension over CUQIpy: ``x0`` may be ``[C, P]``; the sampler then runs C independent chains in one
kernel launch sequence and returns a :class:`BatchedSamples`.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from . import diagnostics
from .engine import (MHEngine, SurrogateSpec, PATH_FP32, PATH_TC, SAMPLER_MH, SAMPLER_CWMH,  # noqa: F401
                     SAMPLER_DA)


# -- geometries ----------------------------------------------------------------------------
class Continuous1D:
    def __init__(self, grid):
        self.grid = np.arange(grid) if np.isscalar(grid) else np.asarray(grid)
        self.par_dim = len(self.grid)


class Discrete:
    def __init__(self, variables):
        self.variables = list(range(variables)) if np.isscalar(variables) else list(variables)
        self.par_dim = len(self.variables)


# -- model -----------------------------------------------------------------------------------
class _ModelOfDistribution:
    """Result of ``cuqi_model(x_distribution)``: the model applied to a random variable."""

    def __init__(self, model, dist):
        self.model, self.dist = model, dist


class Model:
    def __init__(self, forward, jacobian=None, range_geometry=None, domain_geometry=None, gradient=None):
        self.forward_func = forward
        self.jacobian = jacobian
        self.range_geometry = range_geometry
        self.domain_geometry = domain_geometry

    @property
    def domain_dim(self):
        return self.domain_geometry.par_dim

    @property
    def range_dim(self):
        return self.range_geometry.par_dim

    def forward(self, x):
        return self.forward_func(np.asarray(x, np.float64))

    def __call__(self, x):
        if isinstance(x, (Uniform, Gaussian)):
            return _ModelOfDistribution(self, x)
        return self.forward(x)


# -- distributions -----------------------------------------------------------------------------
class Uniform:
    def __init__(self, low, high):
        self.low = np.atleast_1d(np.asarray(low, np.float64))
        self.high = np.atleast_1d(np.asarray(high, np.float64))
        self.dim = self.low.size

    def logpdf(self, x):
        x = np.asarray(x, np.float64)
        if np.any(x < self.low) or np.any(x > self.high):
            return -np.inf
        return float(np.log(1.0 / np.prod(self.high - self.low)))

    logd = logpdf

    def sample(self, N=1, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        s = rng.uniform(self.low, self.high, size=(N, self.dim)).T
        return s[:, 0] if N == 1 else s


class Gaussian:
    """``Gaussian(mean, cov)``: ``cov`` scalar, vector (diagonal) or dense matrix."""

    def __init__(self, mean=None, cov=None):
        self.mean = mean
        self.cov = cov
        self.is_conditional = isinstance(mean, _ModelOfDistribution)
        if not self.is_conditional:
            self.mean = np.atleast_1d(np.asarray(mean, np.float64))
            self.dim = self.mean.size

    def _dense_cov(self, dim):
        c = np.asarray(self.cov, np.float64)
        if c.ndim == 0:
            return np.eye(dim) * float(c)
        if c.ndim == 1:
            return np.diag(c)
        return c

    def iid_variance(self, dim) -> float:
        """sigma^2 if ``cov == sigma^2 * I`` (what the notebooks build), else ValueError."""
        c = self._dense_cov(dim)
        s2 = float(c[0, 0])
        if not np.array_equal(c, np.eye(dim) * s2):
            raise ValueError("the GPU likelihood supports cov = sigma^2 * I only")
        return s2

    def sample(self, N=1, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        c = self._dense_cov(self.dim)
        L = np.linalg.cholesky(c)
        s = self.mean[:, None] + L @ rng.standard_normal((self.dim, N))
        return s[:, 0] if N == 1 else s

    def logd(self, x):
        c = self._dense_cov(self.dim)
        r = np.asarray(x, np.float64) - self.mean
        _, logdet = np.linalg.slogdet(c)
        return float(-0.5 * (logdet + r @ np.linalg.solve(c, r) + self.dim * np.log(2 * np.pi)))

    logpdf = logd


class Posterior:
    """Uniform prior x Gaussian iid likelihood around the surrogate, conditioned on ``y_obs``."""

    def __init__(self, prior: Uniform, likelihood: Gaussian, data):
        if not likelihood.is_conditional:
            raise ValueError("likelihood must be Gaussian(mean=model(x_distribution), cov=...)")
        self.prior = prior
        self.likelihood_dist = likelihood
        self.model: Model = likelihood.mean.model
        self.data = np.asarray(data, np.float64)
        self.dim = prior.dim
        self.sigma2 = likelihood.iid_variance(self.data.size)

    def logd(self, x):
        f = np.asarray(self.model.forward(x), np.float64)
        T = self.data.size
        r = self.data - f
        like = -0.5 * (T * np.log(self.sigma2) + np.sum(r * r) / self.sigma2 + T * np.log(2 * np.pi))
        return like + self.prior.logpdf(x)


class JointDistribution:
    def __init__(self, *dists):
        self.dists = dists

    def __call__(self, *args, **kwargs):
        data = args[0] if args else None
        if kwargs:
            if len(kwargs) != 1:
                raise ValueError("condition on exactly one observed variable")
            data = next(iter(kwargs.values()))
        prior = next((d for d in self.dists if isinstance(d, Uniform)), None)
        like = next((d for d in self.dists if isinstance(d, Gaussian) and d.is_conditional), None)
        if prior is None or like is None or data is None:
            raise ValueError("need a Uniform prior, a conditional Gaussian likelihood and observed data")
        return Posterior(prior, like, data)


# -- samples ---------------------------------------------------------------------------------
class Samples:
    """``cuqi.samples.Samples`` look-alike: ``.samples`` is ``(dim, Ns)`` float64."""

    def __init__(self, samples, loglike_eval=None, acc_rate=None, scale=None, geometry=None):
        self.samples = np.asarray(samples, np.float64)
        self.loglike_eval = loglike_eval
        self.acc_rate = acc_rate
        self.scale = scale
        self.geometry = geometry

    @property
    def shape(self):
        return self.samples.shape

    @property
    def Ns(self):
        return self.samples.shape[-1]

    def burnthin(self, Nb, Nt=1):
        ll = None if self.loglike_eval is None else self.loglike_eval[Nb::Nt]
        return Samples(self.samples[:, Nb::Nt], ll, self.acc_rate, self.scale, self.geometry)

    def mean(self):
        return self.samples.mean(axis=-1)

    def variance(self):
        return self.samples.var(axis=-1)

    def compute_ess(self):
        return np.array([diagnostics.ess_bulk(self.samples[j][None, :]) for j in range(self.samples.shape[0])])

    def compute_rhat(self, chains: Sequence["Samples"]):
        allc = [self] + list(chains)
        if len(allc) < 2:
            raise TypeError("compute_rhat needs at least one other chain")
        if any(c.shape != self.shape for c in allc):
            raise TypeError("all chains must have the same shape")
        return np.array([diagnostics.rhat_rank(np.stack([c.samples[j] for c in allc]))
                         for j in range(self.samples.shape[0])])

    def plot_trace(self, *a, **k):
        raise NotImplementedError("plotting is out of scope (matplotlib/arviz not in this image)")


class BatchedSamples:
    """C chains at once: ``.samples`` is ``(C, dim, Ns)``; indexing yields per-chain :class:`Samples`."""

    def __init__(self, samples, loglike_eval, acc_rate, scale):
        self.samples = np.asarray(samples, np.float64)
        self.loglike_eval = loglike_eval
        self.acc_rate = np.asarray(acc_rate)
        self.scale = np.asarray(scale)

    def __len__(self):
        return self.samples.shape[0]

    def __getitem__(self, c) -> Samples:
        ll = None if self.loglike_eval is None else self.loglike_eval[c]
        return Samples(self.samples[c], ll, float(self.acc_rate[c]), float(self.scale[c]))

    def __iter__(self):
        return (self[c] for c in range(len(self)))

    def burnthin(self, Nb, Nt=1):
        ll = None if self.loglike_eval is None else self.loglike_eval[:, Nb::Nt]
        return BatchedSamples(self.samples[:, :, Nb::Nt], ll, self.acc_rate, self.scale)

    def mean(self):
        return self.samples.mean(axis=(0, 2))

    def std(self):
        return self.samples.transpose(1, 0, 2).reshape(self.samples.shape[1], -1).std(axis=1)

    def _device_samples(self, device):
        import torch
        return torch.as_tensor(np.ascontiguousarray(self.samples.transpose(2, 1, 0))).to(f"cuda:{device}")   # [Ns, P, C]

    def compute_ess(self, on_device: Optional[bool] = None, device: int = 0):
        """Bulk effective sample size per parameter over all chains (arviz's default, rank-normalised).  Large batches
        (or ``on_device=True``) are ranked and reduced on the GPU (``diagnostics_device.ess_bulk``): shipping 10^6
        chains to arviz is the bottleneck the device path removes."""
        C, P, Ns = self.samples.shape
        if on_device is None:
            on_device = C * Ns > 2_000_000
        if not on_device:
            return np.array([diagnostics.ess_bulk(self.samples[:, j, :]) for j in range(P)])
        from . import diagnostics_device
        return diagnostics_device.ess_bulk(self._device_samples(device))

    def compute_rhat(self, on_device: Optional[bool] = None, device: int = 0):
        C, P, Ns = self.samples.shape
        if on_device is None:
            on_device = C * Ns > 2_000_000
        if not on_device:
            return np.array([diagnostics.rhat_rank(self.samples[:, j, :]) for j in range(P)])
        from . import diagnostics_device
        return diagnostics_device.rhat_rank(self._device_samples(device))


# -- sampler ---------------------------------------------------------------------------------
class MH:
    """``cuqi.sampler.MH`` look-alike running on the GPU.

    ``MH(target=posterior, proposal=Gaussian(mean=zeros, cov), x0=x0).sample_adapt(N, Nb)``.
    ``proposal`` must be a zero-mean Gaussian (symmetric, as CUQIpy requires); ``scale=None`` means
    0.1 in ``sample_adapt`` (CUQIpy default) and is an error in ``sample``.
    The reference draws from the unseeded global numpy RNG; here the stream is Philox keyed by ``seed``.
    """

    def __init__(self, target: Posterior, proposal: Optional[Gaussian] = None, scale=None, x0=None, dim=None,
                 seed: int = 0, path: int = PATH_TC, device: int = 0, chain_id0: int = 0):
        if not isinstance(target, Posterior):
            raise TypeError("target must be the Posterior built by JointDistribution(...)(y=...)")
        self.target = target
        self.dim = target.dim if dim is None else dim
        if proposal is None:
            proposal = Gaussian(np.zeros(self.dim), 1.0)
        if not isinstance(proposal, Gaussian) or proposal.is_conditional or np.any(proposal.mean != 0):
            raise ValueError("MH proposal must be a symmetric (zero-mean Gaussian) distribution")
        self.proposal = proposal
        self.scale = scale
        self.x0 = np.ones(self.dim) if x0 is None else np.asarray(x0, np.float64)
        self.seed, self.path, self.device, self.chain_id0 = seed, path, device, chain_id0
        fwd = target.model.forward_func
        if not hasattr(fwd, "spec"):
            raise TypeError("the model's forward function must come from create_forward_model_function")
        self.spec: SurrogateSpec = fwd.spec

    #: sampler variant run by the engine; subclasses override
    _sampler = SAMPLER_MH

    def _engine(self, C) -> MHEngine:
        cov = self.proposal._dense_cov(self.dim)
        eng = MHEngine(self.spec, self.target.data, self.target.sigma2, self.target.prior.low,
                       self.target.prior.high, cov, n_chains=C, path=self.path, device=self.device)
        self._configure(eng)
        return eng

    def _configure(self, eng: MHEngine):
        pass

    def _adapt_window(self, N):
        return int(0.1 * N)

    def sample(self, N, Nb=0):
        if self.scale is None:
            raise ValueError("scale must be set for non-adaptive sampling")
        return self._run(N, Nb, adapt=False)

    def sample_adapt(self, N, Nb=0):
        if self.scale is None:
            self.scale = 0.1
        return self._run(N, Nb, adapt=True)

    def _run(self, N, Nb, adapt):
        x0 = np.atleast_2d(self.x0)
        if x0.shape[1] != self.dim:
            raise ValueError(f"x0 must have {self.dim} entries per chain")
        C = x0.shape[0]
        eng = self._engine(C)
        try:
            eng.init_chains(x0, scale0=float(self.scale), seed=self.seed, chain_id0=self.chain_id0)
            Na = self._adapt_window(N) if adapt else 0
            Ns = N + Nb
            const = eng.loglik_const + eng.logprior_const
            st0 = eng.state()
            pre = []
            if Nb > 0:                                   # samples[:, :Nb] are dropped by CUQIpy
                eng.run(Nb - 1, Na, 1, keep_samples=False) if Nb > 1 else None
                a1 = eng.state()["accept_count"].cpu().numpy().astype(np.int64)
                smp, ll = eng.run(Ns - Nb, Na, 1, keep_samples=True, keep_loglik=True)
                acc_kept = eng.state()["accept_count"].cpu().numpy().astype(np.int64) - a1
            else:
                smp, ll = eng.run(Ns - 1, Na, 1, keep_samples=True, keep_loglik=True)
                pre = [(st0["x"].cpu().numpy(), st0["loglik"].cpu().numpy())]
                per = self.dim if self._sampler == SAMPLER_CWMH else 1                       # CWMH counts per component
                acc_kept = eng.state()["accept_count"].cpu().numpy().astype(np.int64) + per   # acc[0] = 1
            st = eng.state()
            S = np.empty((C, self.dim, N))
            LL = np.empty((C, N))
            k = 0
            for xs, l0 in pre:
                S[:, :, 0] = xs.T
                LL[:, 0] = l0
                k = 1
            if smp is not None:
                S[:, :, k:] = smp.cpu().numpy().transpose(2, 1, 0)
                LL[:, k:] = ll.cpu().numpy().T
            LL = LL + const
            acc_rate = acc_kept / float(N * (self.dim if self._sampler == SAMPLER_CWMH else 1))
            scale = st["scale"].cpu().numpy()
            self.scale = float(scale[0]) if C == 1 else scale
        finally:
            eng.close()
        if np.asarray(self.x0).ndim == 1:
            print("\nAverage acceptance rate:", acc_rate[0], "MCMC scale:", scale[0], "\n")
            return Samples(S[0], LL[0], float(acc_rate[0]), float(scale[0]), self.target.model.domain_geometry)
        return BatchedSamples(S, LL, acc_rate, scale)


class DelayedAcceptanceMH(MH):
    """Two-stage (Christen & Fox 2005) variant of :class:`MH`: the first stage scores every
    ``stride``-th time point of the trace.  Not part of CUQIpy 0.8.0 nor of the reference
    (SURVEY.md F2); same constructor and return types as :class:`MH`."""
    _sampler = SAMPLER_DA

    def __init__(self, *a, stride: int = 4, coarse_model=None, coarse_scaler=None, **k):
        """``coarse_model`` / ``coarse_scaler``: the reference's FFT surrogate (``NN_Model`` loaded from
        ``model_VoltageSignal_fft.keras``) and the scaler of its data processor (``preprocessing(config_I_fft).scaler``);
        when given they replace the time-strided first stage (``stride`` is then ignored)."""
        super().__init__(*a, **k)
        self.stride = int(stride)
        self.coarse_model, self.coarse_scaler = coarse_model, coarse_scaler
        if (coarse_model is None) != (coarse_scaler is None):
            raise ValueError("give both coarse_model and coarse_scaler or neither")

    def _configure(self, eng):
        if self.coarse_model is not None:
            net = getattr(self.coarse_model, "model", self.coarse_model)
            eng.set_coarse_fft(net.layers, self.coarse_scaler.mean_, self.coarse_scaler.scale_)
            eng.set_sampler(SAMPLER_DA, da_stride=0)
        else:
            eng.set_sampler(SAMPLER_DA, da_stride=self.stride)


class CWMH(MH):
    """``cuqi.sampler.CWMH`` look-alike (component-wise MH; never used by the reference).
    ``scale`` is the per-component proposal std (scalar or length-dim vector); the proposal is
    CUQIpy's default ``Normal(location, scale)``; adaptation uses ``Na = int(0.012*N)`` and the target
    ``0.21/dim + 0.23``."""
    _sampler = SAMPLER_CWMH

    def __init__(self, target, proposal=None, scale=1, x0=None, dim=None, **k):
        if proposal is not None:
            raise ValueError("CWMH on the GPU supports CUQIpy's default Normal(location, scale) proposal only")
        super().__init__(target, None, 1.0, x0, dim, **k)
        self.cw_scale = np.broadcast_to(np.asarray(scale, np.float64), (self.dim,)).copy()

    def _configure(self, eng):
        eng.set_sampler(SAMPLER_CWMH, cw_scale=self.cw_scale)

    def _adapt_window(self, N):
        return int(0.012 * N)
