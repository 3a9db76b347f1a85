"""Oracle of the FFT coarse surrogate used as the first stage of delayed acceptance (TEST INFRASTRUCTURE ONLY).

The reference trains ``model_VoltageSignal_fft.keras`` on features of the trace
(src/utils/preprocessing.py:90-103: ``[10*mean, |F_1..7|, 100*angle(F_1..7)]`` with ``F = fft(y - mean)[:8]``) and
reconstructs a band-limited trace from its prediction in
tests/Xaccelerometer_quality_factor/surrogate_fft.ipynb cell 15::

    mean = 0.1*predict[0]; abs_ = predict[1:8]; angle_ = 0.01*predict[8:]
    truncated_fft[1:8]  = abs_*(cos(angle_) + 1j*sin(angle_))
    truncated_fft[-7:]  = (conjugates, mirrored)
    reconstructed = mean + ifft(truncated_fft)

The reference never uses it inside a sampler: "parity unpinned" for the delayed-acceptance use; the
reconstruction itself follows the notebook line by line.
"""
from __future__ import annotations

import numpy as np

from .surrogate import mlp_forward

N_HARM = 7


def features_to_trace(pred, T: int) -> np.ndarray:
    """surrogate_fft.ipynb cell 15, vectorised over leading axes: pred [..., 15] -> trace [..., T] (float64)."""
    pred = np.asarray(pred, np.float64)
    mean = 0.1 * pred[..., :1]
    amp, ang = pred[..., 1:1 + N_HARM], 0.01 * pred[..., 1 + N_HARM:1 + 2 * N_HARM]
    spec = np.zeros(pred.shape[:-1] + (T,), dtype=complex)
    spec[..., 1:1 + N_HARM] = amp * (np.cos(ang) + 1j * np.sin(ang))
    spec[..., -N_HARM:] = (amp * (np.cos(ang) - 1j * np.sin(ang)))[..., ::-1]
    return mean + np.fft.ifft(spec, axis=-1).real


def batched_coarse_forward(layers, scaler_mean, scaler_scale, X, T: int) -> np.ndarray:
    """X [n, P] unscaled parameters -> coarse traces [n, T]: scaler in f64, float32 MLP, reconstruction in f64."""
    rows = ((np.asarray(X, np.float64) - scaler_mean) / scaler_scale).astype(np.float32)
    return features_to_trace(mlp_forward(layers, rows), T)


def coarse_loglik(layers, scaler_mean, scaler_scale, X, y_obs, sigma2, low, high) -> np.ndarray:
    """-0.5*||y_obs - y_c(x)||^2/sigma2 (the device convention: no additive constants), -inf outside the prior box."""
    X = np.atleast_2d(np.asarray(X, np.float64))
    f = batched_coarse_forward(layers, scaler_mean, scaler_scale, X, len(y_obs))
    r = np.asarray(y_obs, np.float64)[None, :] - f
    ll = -0.5 * np.sum(r * r, axis=1) / sigma2
    inb = np.all((X >= low) & (X <= high), axis=1)
    return np.where(inb, ll, -np.inf)
