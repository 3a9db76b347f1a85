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


_ACOV_MAX_DRAWS = 20480        # mems_chain_autocov: draws of one (half-)chain that fit in shared memory


class BatchedSamples:
    """C chains at once: ``.samples`` is ``(C, dim, Ns)``; indexing yields per-chain :class:`Samples`.

    The draws stay where the sampler wrote them -- a device tensor ``[Ns, dim, C]`` (chain-minor, the kernel's coalesced
    layout) -- until somebody asks for the numpy view: ``burnthin`` is a strided view, ``mean`` / ``std`` /
    ``compute_ess`` / ``compute_rhat`` reduce on the device, and ``.samples`` is materialised once (one transposing copy
    on the device, one pinned device-to-host copy).  Built from a host array it behaves exactly as before.
    """

    def __init__(self, samples, loglike_eval, acc_rate, scale, device_samples=None, device_loglike=None):
        self._samples = None if samples is None else np.asarray(samples, np.float64)
        self._loglike = loglike_eval
        self._dev = device_samples              # torch [Ns, dim, C] f64 or None
        self._dev_ll = device_loglike           # torch [Ns, C] f64 (already with the additive constants) or None
        self.acc_rate = np.asarray(acc_rate)
        self.scale = np.asarray(scale)

    # -- lazy host views ---------------------------------------------------------------------
    @staticmethod
    def _to_host(t):
        import torch
        out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        out.copy_(t, non_blocking=False)
        return out.numpy()

    @property
    def samples(self):
        if self._samples is None:
            self._samples = self._to_host(self._dev.permute(2, 1, 0).contiguous())      # (C, dim, Ns)
        return self._samples

    @property
    def loglike_eval(self):
        if self._loglike is None and self._dev_ll is not None:
            self._loglike = self._to_host(self._dev_ll.t().contiguous())                # (C, Ns)
        return self._loglike

    @property
    def shape(self):
        if self._dev is not None:
            Ns, P, C = self._dev.shape
            return (C, P, Ns)
        return self._samples.shape

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, c) -> Samples:
        ll = self.loglike_eval
        return Samples(self.samples[c], None if ll is None else ll[c], float(self.acc_rate[c]), float(self.scale[c]))

    def __iter__(self):
        return (self[c] for c in range(len(self)))

    def burnthin(self, Nb, Nt=1):
        if self._dev is not None:
            ll = None if self._dev_ll is None else self._dev_ll[Nb::Nt]
            return BatchedSamples(None, None, self.acc_rate, self.scale, self._dev[Nb::Nt], ll)
        ll = None if self._loglike is None else self._loglike[:, Nb::Nt]
        return BatchedSamples(self._samples[:, :, Nb::Nt], ll, self.acc_rate, self.scale)

    def mean(self):
        if self._dev is not None:
            return self._dev.mean(dim=(0, 2)).cpu().numpy()
        return self._samples.mean(axis=(0, 2))

    def std(self):
        if self._dev is not None:
            m = self._dev.mean(dim=(0, 2), keepdim=True)
            return ((self._dev - m) ** 2).mean(dim=(0, 2)).sqrt().cpu().numpy()
        return self._samples.transpose(1, 0, 2).reshape(self._samples.shape[1], -1).std(axis=1)

    def device_samples(self, device: int = 0):
        """The draws as a device tensor ``[Ns, dim, C]`` (no copy when the sampler left them there)."""
        import torch
        if self._dev is not None:
            return self._dev if self._dev.is_contiguous() else self._dev.contiguous()
        return torch.as_tensor(np.ascontiguousarray(self._samples.transpose(2, 1, 0))).to(f"cuda:{device}")

    _device_samples = device_samples

    def compute_ess(self, on_device: Optional[bool] = None, device: int = 0, across_ranks: bool = False):
        """Bulk effective sample size per parameter over all chains (arviz's default, rank-normalised).  Large batches,
        device-resident draws (or ``on_device=True``) are ranked and reduced on the GPU by the library's own kernels
        (``diagnostics_device.ess_bulk``): shipping 10^6 chains to arviz is the bottleneck the device path removes
        (SURVEY.md 8f1).  ``across_ranks=True``: this object holds one shard of a ``torch.distributed`` run and the
        estimate is over the chains of every rank (global ranks; same number on any GPU count)."""
        C, P, Ns = self.shape
        if on_device is None:
            on_device = across_ranks or self._dev is not None or C * Ns > 2_000_000
        if on_device and Ns // 2 > _ACOV_MAX_DRAWS:
            # the device autocovariance kernel stages a half-chain in shared memory (<= 20480 draws): longer chains go to the
            # host estimator when that is affordable, otherwise they are thinned to the limit first (ESS of the kept draws)
            if C * Ns <= 50_000_000 and not across_ranks:
                on_device = False
            else:
                return self.burnthin(0, (Ns // 2) // _ACOV_MAX_DRAWS + 1).compute_ess(on_device=True, device=device,
                                                                                        across_ranks=across_ranks)
        if not on_device:
            return np.array([diagnostics.ess_bulk(self.samples[:, j, :]) for j in range(P)])
        from . import diagnostics_device
        return diagnostics_device.ess_bulk(self.device_samples(device), group=None if across_ranks else False)

    def compute_rhat(self, on_device: Optional[bool] = None, device: int = 0, across_ranks: bool = False):
        """Rank-normalised split R-hat per parameter over all chains (arviz's default); see :meth:`compute_ess`."""
        C, P, Ns = self.shape
        if on_device is None:
            on_device = across_ranks or self._dev is not None or C * Ns > 2_000_000
        if not on_device:
            return np.array([diagnostics.rhat_rank(self.samples[:, j, :]) for j in range(P)])
        from . import diagnostics_device
        return diagnostics_device.rhat_rank(self.device_samples(device), group=None if across_ranks else False)


# -- sampler ---------------------------------------------------------------------------------
_ENGINES: dict = {}        # (surrogate, chains, path, device, sampler setup) -> MHEngine, oldest first
_ENGINE_CACHE = 4


def close_engines():
    """Release the cached GPU engines (handles, device buffers)."""
    while _ENGINES:
        _ENGINES.pop(next(iter(_ENGINES))).close()


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
        # x0: [P] (one chain, cuqi), [C, P] (C chains at once); a torch tensor (e.g. pinned host memory) is taken as it is
        self.x0 = np.ones(self.dim) if x0 is None else (x0 if hasattr(x0, "is_cuda") else np.asarray(x0, np.float64))
        self.seed, self.path, self.device, self.chain_id0 = seed, path, device, chain_id0
        fwd = target.model.forward_func
        if not hasattr(fwd, "spec"):
            raise TypeError("the model's forward function must come from create_forward_model_function")
        self.spec: SurrogateSpec = fwd.spec

    #: sampler variant run by the engine; subclasses override
    _sampler = SAMPLER_MH

    def _engine(self, C) -> MHEngine:
        """One engine (handle, device buffers, weight image) per (surrogate, chain count, path, device, sampler setup);
        later calls with the same setup only refresh the problem constants."""
        cov = self.proposal._dense_cov(self.dim)
        key = (id(self.spec), int(C), int(self.path), int(self.device), self._sampler, self._engine_key())
        eng = _ENGINES.get(key)
        args = (self.target.data, self.target.sigma2, self.target.prior.low, self.target.prior.high, cov)
        if eng is None:
            while len(_ENGINES) >= _ENGINE_CACHE:
                _ENGINES.pop(next(iter(_ENGINES))).close()
            eng = MHEngine(self.spec, *args, n_chains=C, path=self.path, device=self.device)
            self._configure(eng)
            eng._problem_args = None
            _ENGINES[key] = eng
        sig = tuple(np.asarray(a, np.float64).tobytes() for a in args)
        if getattr(eng, "_problem_args", None) != sig:
            if eng._problem_args is not None:
                eng.set_problem(*args)
            eng._problem_args = sig
        return eng

    def _engine_key(self):
        return ()

    def _configure(self, eng: MHEngine):
        pass

    def _adapt_window(self, N):
        return int(0.1 * N)

    def sample(self, N, Nb=0, Nt=1):
        if self.scale is None:
            raise ValueError("scale must be set for non-adaptive sampling")
        return self._run(N, Nb, Nt, adapt=False)

    def sample_adapt(self, N, Nb=0, Nt=1):
        """``N`` samples after ``Nb`` burn-in updates, like ``cuqi.sampler.MH.sample_adapt``.  Extension: ``Nt`` keeps
        every ``Nt``-th of them (what ``Samples.burnthin(0, Nt)`` would select) without ever storing the others."""
        if self.scale is None:
            self.scale = 0.1
        return self._run(N, Nb, Nt, adapt=True)

    def _run(self, N, Nb, Nt, adapt):
        import torch
        x0 = np.atleast_2d(self.x0) if not isinstance(self.x0, torch.Tensor) else self.x0.reshape(-1, self.dim)
        if x0.shape[1] != self.dim:
            raise ValueError(f"x0 must have {self.dim} entries per chain")
        if N < 1 or Nb < 0 or Nt < 1:
            raise ValueError("need N >= 1, Nb >= 0, Nt >= 1")
        C = x0.shape[0]
        single = (not isinstance(self.x0, torch.Tensor)) and np.asarray(self.x0).ndim == 1
        eng = self._engine(C)
        scale0 = np.asarray(self.scale, np.float64)
        eng.init_chains(x0, scale0=float(scale0.reshape(-1)[0]), seed=self.seed, chain_id0=self.chain_id0)
        Na = self._adapt_window(N) if adapt else 0
        const = eng.loglik_const + eng.logprior_const
        per = self.dim if self._sampler == SAMPLER_CWMH else 1                       # CWMH counts decisions per component
        # cuqi: samples[:, 0] = x0, then N + Nb - 1 updates; the first Nb samples are dropped.  Kept sample k (k = 0, Nt,
        # 2 Nt, ... < N) is the state after Nb + k updates.
        n_keep = (N - 1) // Nt + 1
        dev = f"cuda:{eng.device}"
        smp = torch.empty(n_keep, self.dim, C, dtype=torch.float64, device=dev)
        ll = torch.empty(n_keep, C, dtype=torch.float64, device=dev)
        if Nb > 0:                                                                    # burn-in: only its last state is kept;
            if Nb > 1:                                                                # cuqi's acc[Nb:] starts with update Nb's decision
                eng.run_into(Nb - 1, Na, Nb, None, None)
            acc0 = eng.accept_count()
            eng.run_into(1, Na, 1, smp[:1], ll[:1])
        else:
            st0 = eng.state()
            smp[0].copy_(st0["x"])
            ll[0].copy_(st0["loglik"])
            acc0 = None
        done = 0
        if n_keep > 1:
            eng.run_into((n_keep - 1) * Nt, Na, Nt, smp[1:], ll[1:])
            done = (n_keep - 1) * Nt
        if N - 1 - done > 0:                                                          # the tail cuqi still samples (and adapts on)
            eng.run_into(N - 1 - done, Na, N, None, None)
        st = eng.state()
        acc_kept = st["accept_count"].to(torch.int64)
        acc_kept = acc_kept - acc0.to(torch.int64) if acc0 is not None else acc_kept + per   # acc[0] = 1 of cuqi
        acc_rate = (acc_kept.double() / float(N * per)).cpu().numpy()
        if self._sampler == SAMPLER_CWMH:
            scale = eng.aux_state()["scale_cw"].t().contiguous().cpu().numpy()                # [C, dim], cuqi keeps the vector
            self.cw_scale = scale[0].copy()
        else:
            scale = st["scale"].cpu().numpy()
        self.scale = (float(scale[0]) if scale.ndim == 1 else scale[0]) if C == 1 else scale
        ll += const
        if single:
            print("\nAverage acceptance rate:", acc_rate[0], "MCMC scale:", scale[0], "\n")
            return Samples(smp[:, :, 0].t().cpu().numpy(), ll[:, 0].cpu().numpy(), float(acc_rate[0]),
                           scale[0] if scale.ndim > 1 else float(scale[0]), self.target.model.domain_geometry)
        return BatchedSamples(None, None, acc_rate, scale, device_samples=smp, device_loglike=ll)


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

    def _engine_key(self):
        return ("da", self.stride, id(self.coarse_model))

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

    def _engine_key(self):
        return ("cw", tuple(float(v) for v in self.cw_scale))

    def _configure(self, eng):
        eng.set_sampler(SAMPLER_CWMH, cw_scale=self.cw_scale)

    def _adapt_window(self, N):
        return int(0.012 * N)
