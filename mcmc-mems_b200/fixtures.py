"""Inverse-problem fixtures: everything the reference's notebook builds before it calls the sampler
(tests/Xaccelerometer_geometric/identification.ipynb cells 6-12), frozen into one ``.npz`` by
``tests/golden/make_golden.py`` (which imports the reference's own ``src/utils/preprocessing.py`` and reads its
``models/*.keras`` files): surrogate weights, time grid, fitted ``StandardScaler`` constants, observation, prior box,
least-squares optima and the proposal covariance.  The GPU box has no copy of the reference tree, so ``bench.py`` and
``__graft_entry__.smoke()`` rebuild the reference-shaped objects from these files.
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np

from .model import NN_Model
from .preprocessing import FixtureProcessor

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fixture_path(name: str) -> str:
    """``'geometric'`` / ``'qfactor'`` -> the committed fixture file."""
    return os.path.join(_ROOT, "tests", "golden", f"{name}.npz")


def load_fixture(name_or_path: str, n_time: int = 0):
    """-> namespace with ``processor``, ``nn_model`` (an :class:`NN_Model` with the fixture's weights), ``y_obs``,
    ``sigma2``, ``low``, ``high``, ``cov`` (proposal covariance = inv(J^T J) at the LSQ optimum), ``x_opts`` (LSQ optima
    of the notebook's start points), ``x_true``.

    ``n_time`` > 0 resamples the time grid to that many points on the same interval (the synthetic long-trace
    configuration of BASELINE.json) and draws a new observation around the surrogate's own prediction at ``x_true``
    (seeded), evaluated on the GPU through the forward closure.
    """
    path = name_or_path if os.path.exists(name_or_path) else fixture_path(name_or_path)
    d = np.load(path)
    layers = [(d[f"W{i}"], d[f"b{i}"], str(d["activations"][i])) for i in range(int(d["n_layers"]))]
    nn = NN_Model()
    nn.set_layers(layers)
    fx = SimpleNamespace(
        name=os.path.splitext(os.path.basename(path))[0], layers=layers, nn_model=nn,
        processor=FixtureProcessor(d["time"], d["scaler_mean"], d["scaler_scale"]),
        y_obs=np.asarray(d["y_obs"], np.float64), sigma2=float(d["sigma2"]), low=np.asarray(d["low"], np.float64),
        high=np.asarray(d["high"], np.float64), cov=np.asarray(d["cov"], np.float64), x_opts=np.asarray(d["x_opts"], np.float64),
        x_true=np.asarray(d["x_true"], np.float64))
    if n_time and n_time != len(fx.processor.time):
        from .utils import create_forward_model_function
        t = fx.processor.time
        fx.processor = FixtureProcessor(np.linspace(t[0], t[-1], n_time), d["scaler_mean"], d["scaler_scale"])
        f = create_forward_model_function(fx.processor, nn)(fx.x_true)
        fx.y_obs = np.asarray(f, np.float64) + np.sqrt(fx.sigma2) * np.random.default_rng(0).standard_normal(n_time)
    return fx


def posterior_of(fx, forward_model=None):
    """The CUQIpy-shaped posterior of the fixture, built the way the notebook does (cells 8, 10, 14)."""
    from . import cuqi_compat as cq
    from .utils import create_forward_model_function
    fwd = forward_model or create_forward_model_function(fx.processor, fx.nn_model)
    T, P = len(fx.processor.time), len(fx.low)
    model = cq.Model(forward=fwd, range_geometry=cq.Continuous1D(T), domain_geometry=cq.Discrete(P))
    x = cq.Uniform(fx.low, fx.high)
    y = cq.Gaussian(mean=model(x), cov=fx.sigma2 * np.eye(T))
    return cq.JointDistribution(x, y)(y=fx.y_obs)
