"""Thin Python owner of one ``mems_handle_t``: torch tensors hold every device buffer, ctypes
passes their pointers through the C ABI (include/mems_b200.h).  No arithmetic happens here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import MemsError, PATH_FP32, PATH_TC, SAMPLER_MH, SAMPLER_CWMH, SAMPLER_DA  # noqa: F401


def _require_cuda(device: int = 0):
    if not torch.cuda.is_available():
        raise MemsError("CUDA is not available: the surrogate-MH path runs on sm_100a only "
                        "(no CPU fallback)")
    if device >= torch.cuda.device_count():
        raise MemsError(f"CUDA device {device} not present")


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _stream(device: int):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _host(a, dtype=np.float64):
    return np.ascontiguousarray(np.asarray(a, dtype=dtype))


@dataclass
class SurrogateSpec:
    """Everything the kernels need to know about one inverse problem."""
    layers: Sequence            # [(kernel[in,out] f32, bias f32, activation)] * 7
    time: np.ndarray            # [T]
    scaler_mean: np.ndarray     # [P+1]
    scaler_scale: np.ndarray    # [P+1]

    @property
    def n_params(self) -> int:
        return int(np.asarray(self.layers[0][0]).shape[0]) - 1

    @property
    def n_time(self) -> int:
        return int(len(self.time))

    def validate(self):
        if len(self.layers) != _lib.MEMS_HIDDEN + 1:
            raise ValueError(f"expected {_lib.MEMS_HIDDEN + 1} Dense layers, got {len(self.layers)}")
        P = self.n_params
        if not 1 <= P <= _lib.MEMS_PMAX:
            raise ValueError(f"n_params={P} outside [1,{_lib.MEMS_PMAX}]")
        shapes = [(P + 1, 64)] + [(64, 64)] * 5 + [(64, 1)]
        for i, ((k, b, act), shp) in enumerate(zip(self.layers, shapes)):
            if tuple(np.asarray(k).shape) != shp or tuple(np.asarray(b).shape) != (shp[1],):
                raise ValueError(f"layer {i}: kernel {np.asarray(k).shape} / bias {np.asarray(b).shape}, expected {shp}")
            want = "linear" if i == 6 else "tanh"
            if act != want:
                raise ValueError(f"layer {i}: activation {act!r}, expected {want!r}")
        if len(self.scaler_mean) != P + 1 or len(self.scaler_scale) != P + 1:
            raise ValueError("scaler constants must have P+1 entries")


class MHEngine:
    """Batched random-walk Metropolis-Hastings over ``n_chains`` independent chains on one GPU.

    Mirrors, in batch, what ``cuqi.sampler.MH`` does for one chain in the reference
    (tests/Xaccelerometer_geometric/identification.ipynb:300-302).
    """

    def __init__(self, spec: SurrogateSpec, y_obs, sigma2: float, low, high, cov,
                 n_chains: int, path: int = PATH_TC, device: int = 0, chains_per_block: int = 0):
        _require_cuda(device)
        spec.validate()
        self.spec = spec
        self.P, self.T, self.C = spec.n_params, spec.n_time, int(n_chains)
        self.device = int(device)
        self.path = int(path)
        self.sigma2 = float(sigma2)
        self._L = _lib.lib()
        cfg = _lib.MemsConfig(self.P, self.T, self.C, self.path, self.device, int(chains_per_block))
        h = C.c_void_p()
        _lib.check(self._L.mems_create(C.byref(cfg), C.byref(h)), "mems_create")
        self._h = h
        self.sampler = SAMPLER_MH
        self.n_sub = 1
        try:
            Ws = [_host(k, np.float32) for k, _, _ in spec.layers]
            bs = [_host(b, np.float32) for _, b, _ in spec.layers]
            Wp = (C.c_void_p * 7)(*[w.ctypes.data for w in Ws])
            bp = (C.c_void_p * 7)(*[b.ctypes.data for b in bs])
            _lib.check(self._L.mems_set_weights(h, Wp, bp), "mems_set_weights")
            self.set_problem(y_obs, sigma2, low, high, cov)
        except Exception:
            self.close()
            raise

    # -- lifetime ------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.mems_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- constants -----------------------------------------------------------
    def set_problem(self, y_obs, sigma2, low, high, cov):
        P, T = self.P, self.T
        y = _host(y_obs)
        lo, hi = _host(low), _host(high)
        cov = _host(cov).reshape(P, P)
        if y.shape != (T,) or lo.shape != (P,) or hi.shape != (P,):
            raise ValueError("y_obs/low/high have the wrong shape")
        if not np.allclose(cov, cov.T, rtol=1e-10, atol=0):
            raise ValueError("proposal covariance must be symmetric")   # cuqi raises ValueError too
        chol = self._chol_psd(cov)
        t, m, s = _host(self.spec.time), _host(self.spec.scaler_mean), _host(self.spec.scaler_scale)
        self.sigma2 = float(sigma2)
        self.y_obs, self.low, self.high, self.cov, self.chol = y, lo, hi, cov, chol
        _lib.check(self._L.mems_set_problem(self._h, t.ctypes.data, m.ctypes.data, s.ctypes.data,
                                            y.ctypes.data, float(sigma2), lo.ctypes.data,
                                            hi.ctypes.data, chol.ctypes.data), "mems_set_problem")

    @staticmethod
    def _chol_psd(cov):
        """Lower Cholesky factor; zero rows/columns (fixed parameters) are allowed."""
        d = np.diag(cov)
        free = d > 0
        L = np.zeros_like(cov)
        if free.any():
            idx = np.where(free)[0]
            L[np.ix_(idx, idx)] = np.linalg.cholesky(cov[np.ix_(idx, idx)])
        return np.ascontiguousarray(L)

    @property
    def loglik_const(self) -> float:
        """Additive constant dropped on the device: -0.5*T*(log sigma2 + log 2pi)."""
        return -0.5 * self.T * (np.log(self.sigma2) + np.log(2.0 * np.pi))

    @property
    def logprior_const(self) -> float:
        return float(np.log(1.0 / np.prod(self.high - self.low)))

    # -- device helpers --------------------------------------------------------
    def _dev(self, a, dtype=torch.float64):
        if isinstance(a, torch.Tensor):
            return a.to(device=f"cuda:{self.device}", dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(f"cuda:{self.device}")

    def _new(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=f"cuda:{self.device}")

    # -- forward / log-likelihood -----------------------------------------------
    def forward(self, x, want_y=True, want_loglik=True):
        """x: [n, P] (host or device) -> (y [n, T] f32 or None, loglik [n] f64 or None) on device."""
        xd = self._dev(x).reshape(-1, self.P)
        n = xd.shape[0]
        xt = xd.t().contiguous()                       # [P][n]
        y = self._new(n, self.T, dtype=torch.float32) if want_y else None
        ll = self._new(n) if want_loglik else None
        _lib.check(self._L.mems_forward(self._h, _ptr(xt), n, _ptr(y), _ptr(ll), _stream(self.device)), "mems_forward")
        return y, ll

    # -- least-squares initialisation ---------------------------------------------
    def least_squares(self, starts, y_obs=None, bounds=None, max_iter: int = 60, rel_step: float = 0.0,
                      xtol: float = 1e-8, ftol: float = 1e-8):
        """Fit every start point at once (``mems_lsq``).  starts: [n, P]; y_obs: None (the problem's trace),
        [T] or [n, T]; bounds: None (prior box) or (low[P], high[P]); rel_step: finite-difference step relative
        to max(1,|x|) (0 = scipy's eps^(1/3), the reference's setting; ~3e-4 removes the float32 noise from J).
        Returns device tensors (x_opt [n, P], cov [n, P, P] = inv(J^T J), cost [n], iters [n])."""
        xd = self._dev(starts).reshape(-1, self.P)
        n = xd.shape[0]
        xt = xd.t().contiguous()
        yd, n_obs = None, 1
        if y_obs is not None:
            yd = self._dev(y_obs).reshape(-1, self.T).contiguous()
            n_obs = yd.shape[0]
            if n_obs not in (1, n):
                raise ValueError("y_obs must be [T] or [n, T]")
        lo = hi = None
        if bounds is not None:
            lo, hi = _host(bounds[0]), _host(bounds[1])
            if lo.shape != (self.P,) or hi.shape != (self.P,):
                raise ValueError("bounds must be (low[P], high[P])")
        xo = self._new(self.P, n)
        cov = self._new(n, self.P, self.P)
        cost = self._new(n)
        its = self._new(n, dtype=torch.int32)
        _lib.check(self._L.mems_lsq(self._h, _ptr(xt), n, _ptr(yd), n_obs,
                                    None if lo is None else lo.ctypes.data, None if hi is None else hi.ctypes.data,
                                    int(max_iter), float(rel_step), float(xtol), float(ftol), _ptr(xo), _ptr(cov), _ptr(cost), _ptr(its),
                                    _stream(self.device)), "mems_lsq")
        return xo.t().contiguous(), cov, cost, its

    # -- coarse first-stage model of delayed acceptance -----------------------------
    def set_coarse_fft(self, layers, scaler_mean, scaler_scale):
        """Install the reference's FFT surrogate ([P] -> 64x6 tanh -> 15 features) as the DA first stage
        (``mems_set_coarse_fft``); select it with ``set_sampler(SAMPLER_DA, da_stride=0)``."""
        shapes = [(self.P, 64)] + [(64, 64)] * 5 + [(64, 15)]
        if len(layers) != 7:
            raise ValueError(f"expected 7 Dense layers, got {len(layers)}")
        for i, ((k, b, act), shp) in enumerate(zip(layers, shapes)):
            if tuple(np.asarray(k).shape) != shp or tuple(np.asarray(b).shape) != (shp[1],):
                raise ValueError(f"coarse layer {i}: kernel {np.asarray(k).shape} / bias {np.asarray(b).shape}, expected {shp}")
            if act != ("linear" if i == 6 else "tanh"):
                raise ValueError(f"coarse layer {i}: activation {act!r}")
        Ws = [_host(k, np.float32) for k, _, _ in layers]
        bs = [_host(b, np.float32) for _, b, _ in layers]
        Wp = (C.c_void_p * 7)(*[w.ctypes.data for w in Ws])
        bp = (C.c_void_p * 7)(*[b.ctypes.data for b in bs])
        m, sc = _host(scaler_mean), _host(scaler_scale)
        if m.shape != (self.P,) or sc.shape != (self.P,):
            raise ValueError("coarse scaler constants must have P entries")
        _lib.check(self._L.mems_set_coarse_fft(self._h, Wp, bp, m.ctypes.data, sc.ctypes.data), "mems_set_coarse_fft")

    # -- sampler variant ---------------------------------------------------------
    def set_sampler(self, sampler: int, cw_scale=None, da_stride: int = 4):
        """SAMPLER_MH (reference), SAMPLER_CWMH (component-wise; ``cw_scale[P]`` = initial per-component
        proposal std) or SAMPLER_DA (delayed acceptance; coarse stage = every ``da_stride``-th time point, or the
        FFT surrogate installed with :meth:`set_coarse_fft` when ``da_stride == 0``).  Call before :meth:`init_chains`."""
        cw = None if cw_scale is None else _host(np.broadcast_to(np.asarray(cw_scale, np.float64), (self.P,)))
        _lib.check(self._L.mems_set_sampler(self._h, int(sampler), None if cw is None else cw.ctypes.data,
                                            int(da_stride)), "mems_set_sampler")
        self.sampler = int(sampler)
        self.n_sub = self.P if sampler == SAMPLER_CWMH else (2 if sampler == SAMPLER_DA else 1)
        self.da_stride = int(da_stride)

    # -- chains ------------------------------------------------------------------
    def init_chains(self, x0, scale0: float = 0.1, seed: int = 0, chain_id0: int = 0):
        xd = self._dev(x0).reshape(self.C, self.P)
        xt = xd.t().contiguous()
        _lib.check(self._L.mems_init_chains(self._h, _ptr(xt), float(scale0), C.c_uint64(seed),
                                            C.c_uint64(chain_id0), _stream(self.device)), "mems_init_chains")

    def run(self, n_steps: int, adapt_window: int = 0, thin: int = 1, keep_samples=True, keep_loglik=False):
        """Advance every chain by ``n_steps`` updates.  Returns (samples [n_keep,P,C], loglik [n_keep,C])."""
        n_keep = n_steps // thin
        smp = self._new(n_keep, self.P, self.C) if keep_samples and n_keep else None
        ll = self._new(n_keep, self.C) if keep_loglik and n_keep else None
        _lib.check(self._L.mems_run(self._h, int(n_steps), int(adapt_window), int(thin), _ptr(smp), _ptr(ll),
                                    _stream(self.device)), "mems_run")
        return smp, ll

    def run_into(self, n_steps: int, adapt_window: int, thin: int, samples: Optional[torch.Tensor],
                 loglik: Optional[torch.Tensor] = None):
        """Same as :meth:`run` with caller-owned output tensors (no allocation in the timed path)."""
        _lib.check(self._L.mems_run(self._h, int(n_steps), int(adapt_window), int(thin), _ptr(samples),
                                    _ptr(loglik), _stream(self.device)), "mems_run")

    def run_explicit(self, xi, log_u, adapt_window: int = 0):
        """xi: [n_steps, P, C]; log_u: [n_steps, C] (MH) or [n_steps, n_sub, C] (CWMH: n_sub = P, DA: 2).
        Returns (samples [n_steps,P,C], loglik_prop, decisions) with the shape of log_u."""
        xid, lud = self._dev(xi), self._dev(log_u)
        n_steps = xid.shape[0]
        lshape = (n_steps, self.C) if (self.n_sub == 1 and lud.ndim == 2) else (n_steps, self.n_sub, self.C)
        if xid.shape != (n_steps, self.P, self.C) or tuple(lud.shape) != lshape:
            raise ValueError(f"xi must be [n_steps,P,C] and log_u {list(lshape)}")
        smp = self._new(n_steps, self.P, self.C)
        llp = self._new(*lshape)
        dec = self._new(*lshape, dtype=torch.uint8)
        _lib.check(self._L.mems_run_explicit(self._h, n_steps, int(adapt_window), _ptr(xid), _ptr(lud),
                                             _ptr(smp), _ptr(llp), _ptr(dec), _stream(self.device)), "mems_run_explicit")
        return smp, llp, dec

    def draw(self, step0: int, n_steps: int):
        xi = self._new(n_steps, self.P, self.C)
        lu = self._new(n_steps, self.C) if self.n_sub == 1 else self._new(n_steps, self.n_sub, self.C)
        _lib.check(self._L.mems_draw(self._h, C.c_uint64(step0), int(n_steps), _ptr(xi), _ptr(lu),
                                     _stream(self.device)), "mems_draw")
        return xi, lu

    def state(self):
        x = self._new(self.P, self.C)
        ll = self._new(self.C)
        sc = self._new(self.C)
        ac = self._new(self.C, dtype=torch.int32)
        steps = C.c_uint64(0)
        _lib.check(self._L.mems_get_state(self._h, _ptr(x), _ptr(ll), _ptr(sc), _ptr(ac), C.byref(steps),
                                          _stream(self.device)), "mems_get_state")
        return {"x": x, "loglik": ll, "scale": sc, "accept_count": ac, "steps_done": int(steps.value)}

    def accept_count(self):
        """Accepted updates per chain since ``init_chains`` (device tensor, int32)."""
        ac = self._new(self.C, dtype=torch.int32)
        _lib.check(self._L.mems_get_state(self._h, None, None, None, _ptr(ac), None, _stream(self.device)), "mems_get_state")
        return ac

    def aux_state(self):
        """Variant-specific state: per-component scales (CWMH), coarse log-likelihood and promotions (DA)."""
        sc = self._new(self.P, self.C)
        llc = self._new(self.C)
        pc = self._new(self.C, dtype=torch.int32)
        _lib.check(self._L.mems_get_aux(self._h, _ptr(sc), _ptr(llc), _ptr(pc), _stream(self.device)), "mems_get_aux")
        return {"scale_cw": sc, "loglik_coarse": llc, "promote_count": pc}

    def rows_evaluated(self) -> int:
        """Surrogate rows (chain, time point) evaluated by ``run`` launches since ``init_chains`` (synchronises)."""
        v = C.c_uint64(0)
        _lib.check(self._L.mems_rows_evaluated(self._h, C.byref(v), _stream(self.device)), "mems_rows_evaluated")
        return int(v.value)

    @property
    def launches(self) -> int:
        return int(self._L.mems_launch_count(self._h))


def chain_autocov(samples: torch.Tensor, max_lag: int) -> torch.Tensor:
    """samples [n_keep, P, C] f64 on a CUDA device -> [P, max_lag + 2] sums over chains (``mems_chain_autocov``)."""
    if not samples.is_cuda:
        raise MemsError("chain_autocov needs device samples (no CPU fallback)")
    smp = samples.contiguous().double()
    n_keep, P, Cn = smp.shape
    out = torch.empty(P, max_lag + 2, dtype=torch.float64, device=smp.device)
    dev = smp.device.index or 0
    _lib.check(_lib.lib().mems_chain_autocov(dev, _ptr(smp), n_keep, P, Cn, int(max_lag), _ptr(out), _stream(dev)),
               "mems_chain_autocov")
    return out


def mlp_forward(layers, X, device: int = 0):
    """Generic Dense-stack inference on the GPU (NN_Model.predict).  X: [n, in] -> f32 [n, out]."""
    _require_cuda(device)
    L = _lib.lib()
    Ws = [_host(k, np.float32) for k, _, _ in layers]
    bs = [_host(b, np.float32) for _, b, _ in layers]
    for i, (_, _, act) in enumerate(layers):
        want = "linear" if i == len(layers) - 1 else "tanh"
        if act != want:
            raise ValueError(f"layer {i}: activation {act!r} unsupported (need {want!r})")
    nl = len(layers)
    widths = (C.c_int32 * (nl + 1))(Ws[0].shape[0], *[w.shape[1] for w in Ws])
    Wp = (C.c_void_p * nl)(*[w.ctypes.data for w in Ws])
    bp = (C.c_void_p * nl)(*[b.ctypes.data for b in bs])
    Xd = torch.as_tensor(np.ascontiguousarray(np.asarray(X, np.float32)) if not isinstance(X, torch.Tensor) else X,
                         dtype=torch.float32).to(f"cuda:{device}").contiguous()
    if Xd.ndim != 2 or Xd.shape[1] != Ws[0].shape[0]:
        raise ValueError(f"X must be [n,{Ws[0].shape[0]}]")
    n = Xd.shape[0]
    Y = torch.empty(n, Ws[-1].shape[1], dtype=torch.float32, device=Xd.device)
    _lib.check(L.mems_mlp_forward(device, nl, widths, Wp, bp, _ptr(Xd), n, _ptr(Y), _stream(device)), "mems_mlp_forward")
    return Y
