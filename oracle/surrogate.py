"""Oracle: surrogate MLP and forward-model closure (numpy).  Test infrastructure only."""
from __future__ import annotations

import numpy as np


def mlp_forward(layers, X, dtype=np.float32):
    """Keras ``Sequential`` of ``Dense`` layers: ``a <- act(a @ kernel + bias)``.

    Follows src/SurrogateModeling/model.py:56-59 (Dense(units, activation) with
    kernel layout [in, out], ``use_bias=True``) as TF evaluates it: the input is
    cast to the model dtype (float32) first, every layer computes in that dtype.

    layers: sequence of (kernel[in,out], bias[out], activation) triples.
    """
    a = np.asarray(X).astype(dtype)
    for kernel, bias, act in layers:
        a = a @ np.asarray(kernel, dtype) + np.asarray(bias, dtype)
        if act == "tanh":
            a = np.tanh(a)
        elif act != "linear":
            raise ValueError(f"unsupported activation {act!r}")
    return a


def make_forward_model(time, scaler_mean, scaler_scale, layers, dtype=np.float32):
    """Restates ``create_forward_model_function`` (src/InverseProblems/utils.py:22-42).

    ``x[P]`` is tiled to ``[T,P]`` (utils.py:32), the time column is appended
    (utils.py:35), ``StandardScaler.transform`` = ``(X - mean_) / scale_`` is
    applied in float64 (utils.py:38; src/utils/preprocessing.py:170-179), the
    Keras model is called (utils.py:39) and the result flattened to ``dtype[T]``.
    """
    time = np.asarray(time, np.float64)
    T = len(time)
    t_col = time.reshape(-1, 1)
    mean = np.asarray(scaler_mean, np.float64)
    scale = np.asarray(scaler_scale, np.float64)

    def forward_model(x):
        x_rep = np.tile(np.asarray(x, np.float64), (T, 1))
        x_comb = np.hstack((x_rep, t_col))
        scaled = (x_comb - mean) / scale
        return mlp_forward(layers, scaled, dtype).flatten()

    return forward_model


def batched_forward(time, scaler_mean, scaler_scale, layers, X, dtype=np.float32):
    """Same arithmetic as :func:`make_forward_model` for ``X[n,P]`` at once -> ``[n,T]``."""
    time = np.asarray(time, np.float64)
    X = np.atleast_2d(np.asarray(X, np.float64))
    n, P = X.shape
    T = len(time)
    rows = np.empty((n, T, P + 1), np.float64)
    rows[:, :, :P] = X[:, None, :]
    rows[:, :, P] = time[None, :]
    scaled = (rows - np.asarray(scaler_mean, np.float64)) / np.asarray(scaler_scale, np.float64)
    return mlp_forward(layers, scaled.reshape(n * T, P + 1), dtype).reshape(n, T)
