"""Inverse-problem glue with the reference's function names
(src/InverseProblems/utils.py:22-42,83-93): ``create_forward_model_function`` and
``least_squares_optimization``.  The forward model is evaluated on the GPU through the C ABI.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import least_squares

from .engine import MHEngine, SurrogateSpec, PATH_FP32, PATH_TC  # noqa: F401


def surrogate_spec(data_processor, nn_model) -> SurrogateSpec:
    """Collect what the closure at utils.py:26-28 captures: time grid, scaler, Keras layers."""
    model = getattr(nn_model, "model", nn_model)
    layers = model.layers
    if layers and not isinstance(layers[0], tuple):      # tolerate h5lite.DenseLayer lists
        layers = [(l.kernel, l.bias, l.activation) for l in layers]
    sc = data_processor.scaler
    return SurrogateSpec(layers=layers, time=np.asarray(data_processor.time, np.float64),
                         scaler_mean=np.asarray(sc.mean_, np.float64),
                         scaler_scale=np.asarray(sc.scale_, np.float64))


def create_forward_model_function(data_processor, nn_model, path: int = PATH_TC, device: int = 0):
    """``x[P] -> float32[T]`` like the reference closure; also accepts ``x[n,P] -> [n,T]``.  Evaluated on the product
    (tensor-core) path unless ``path=PATH_FP32`` asks for the FP32 check path.

    The returned callable exposes ``.spec`` (weights, scaler constants, time grid) so that the
    batched sampler can find the surrogate without introspecting a closure.
    """
    spec = surrogate_spec(data_processor, nn_model)
    spec.validate()
    P, T = spec.n_params, spec.n_time
    state = {}

    def _engine():
        if "eng" not in state:   # forward only: the likelihood / prior constants are placeholders
            state["eng"] = MHEngine(spec, np.zeros(T), 1.0, np.full(P, -1e300), np.full(P, 1e300),
                                    np.eye(P), n_chains=1, path=path, device=device)
        return state["eng"]

    def forward_model(x):
        x = np.asarray(x, np.float64)
        y, _ = _engine().forward(x.reshape(-1, P), want_y=True, want_loglik=False)
        out = y.cpu().numpy()
        return out.reshape(-1) if x.ndim == 1 else out

    forward_model.spec = spec
    forward_model.weights = spec.layers
    forward_model.scaler_mean, forward_model.scaler_scale, forward_model.time = spec.scaler_mean, spec.scaler_scale, spec.time
    return forward_model


def objective_function(input, exact_outputs, forward_model):
    return forward_model(input).reshape(exact_outputs.shape) - exact_outputs


def least_squares_optimization(y_observed, forward_model, start_point, bounds):
    """``scipy.optimize.least_squares(jac='3-point', bounds)``; returns ``(x_opt, inv(J^T J))``
    exactly as utils.py:83-88 does (the second value is the proposal covariance of the notebooks)."""
    res = least_squares(objective_function, start_point, args=(np.asarray(y_observed), forward_model),
                        bounds=bounds, jac="3-point")
    return res.x, np.linalg.inv(res.jac.T @ res.jac)


def batched_least_squares(y_observed, forward_model, start_points, bounds, sigma2: float = 1.0, path: int = PATH_FP32,
                          device: int = 0, max_iter: int = 60, rel_step: float = 0.0):
    """All start points (and optionally one observed trace per start) fitted together on the GPU.

    Same contract as calling :func:`least_squares_optimization` in a loop, as the notebooks do
    (tests/Xaccelerometer_geometric/identification.ipynb:222-227): returns ``(x_opts [n,P], covs [n,P,P])`` with
    ``covs[i] = inv(J^T J)`` at the optimum of start ``i``.  ``y_observed`` is ``[T]`` or ``[n,T]``.
    """
    spec = forward_model.spec
    P, T = spec.n_params, spec.n_time
    starts = np.asarray(start_points, np.float64).reshape(-1, P)
    y = np.asarray(y_observed, np.float64).reshape(-1, T)
    lo, hi = np.asarray(bounds[0], np.float64), np.asarray(bounds[1], np.float64)
    eng = MHEngine(spec, y[0], sigma2, lo, hi, np.eye(P), n_chains=1, path=path, device=device)
    try:
        x, cov, _, _ = eng.least_squares(starts, y_obs=None if y.shape[0] == 1 else y, bounds=(lo, hi), max_iter=max_iter,
                                           rel_step=rel_step)
        return x.cpu().numpy(), cov.cpu().numpy()
    finally:
        eng.close()


def sensitivity_push_forward(sample, model, scaler, S_avg: float = 4.44):
    """Posterior push-forward through the sensitivity surrogate -- the computation inside
    ``plot_sensitivity_histogram`` (src/InverseProblems/plotsPaper.py:201-210) without the plot.

    ``sample``: ``(P, Ns)`` like ``Samples.samples`` (or ``(C, P, Ns)`` for batched chains); ``model``: an
    ``NN_Model`` / ``DenseStack`` (inference on the GPU); ``scaler``: the fitted scaler of the sensitivity data set.
    Returns ``dict(values, mean, ci_lower, ci_upper)`` with the 5 % / 95 % percentiles the reference shades.
    """
    a = np.asarray(getattr(sample, "samples", sample), np.float64)
    if a.ndim == 3:
        a = np.moveaxis(a, 1, 0).reshape(a.shape[1], -1)          # (C, P, Ns) -> (P, C*Ns)
    rows = scaler.transform(a.T)
    net = getattr(model, "model", model)
    values = np.asarray(net.predict(rows)).reshape(-1) / S_avg
    lo, hi = np.percentile(values, [5, 95])
    return {"values": values, "mean": float(values.mean()), "ci_lower": float(lo), "ci_upper": float(hi)}
