"""Data processor with the attribute surface of the reference's ``preprocessing`` class
(src/utils/preprocessing.py:15-24): ``.config .time .scaler .X_train .X_test .y_train .y_test
.X_train_scaled .X_test_scaled``, ``scale_new_data`` and ``stack_data``.  Only the constants
(``time``, ``scaler.mean_``, ``scaler.scale_``) reach the hot path; this module exists so the
notebook cells that build them run unchanged.  Written from the behaviour, not the source:
same JSON keys, same CSV layout (params, ts, tf, values), same sklearn split and scaler.
"""
from __future__ import annotations

import json

import numpy as np
import pandas as pd
from sklearn.model_selection import train_test_split
from sklearn.preprocessing import MinMaxScaler, StandardScaler


class preprocessing:  # noqa: N801  (reference class name)
    def __init__(self, config_path: str):
        with open(config_path) as fh:
            self.config = json.load(fh)
        self.time = None
        self.scaler = None
        self._read_table()
        self._split()
        self._fit_scaler()

    # -- CSV -> frame ---------------------------------------------------------
    def _read_table(self):
        cfg = self.config
        frame = pd.read_csv(cfg["DATASET_PATH"], header=None, comment="#", dtype=str,
                            float_precision="high").astype(np.float64)
        names = list(cfg["INPUT_COLS"])
        self._kind = cfg["CONFIGURATION"]
        if self._kind in ("I", "fft"):
            dt, t_end = frame.iloc[0, len(names)], frame.iloc[0, len(names) + 1]
            self.time = np.arange(0, t_end, dt)
            n_val = int(t_end / dt) + 1
            frame.columns = names + ["ts", "tf"] + [f"Time={1e3 * dt * i:.2f}ms" for i in range(n_val)]
        else:
            frame.columns = names + ["sensitivity"]
        self.data = frame

    # -- split ------------------------------------------------------------------
    def _split(self):
        cfg = self.config
        names = list(cfg["INPUT_COLS"])
        cols = self.data.columns
        targets = cols[len(names) + 2:-1] if len(cols) > 5 else cols[-1]
        X = self.data[names].values
        y = cfg["Y_SCALING_FACTOR"] * self.data[targets].values
        kw = dict(test_size=cfg["TEST_SIZE"], random_state=cfg["RANDOM_STATE"])
        if self._kind == "fft":
            _, _, self.y_series_train, self.y_series_test = train_test_split(X, y, **kw)
            centred = y - y.mean(axis=1, keepdims=True)
            spec = np.fft.fft(centred)[:, :8]
            y = np.hstack((10 * y.mean(axis=1, keepdims=True), np.abs(spec)[:, 1:], 100 * np.angle(spec)[:, 1:]))
        self.X_train, self.X_test, self.y_train, self.y_test = train_test_split(X, y, **kw)

    # -- scaling ------------------------------------------------------------------
    @staticmethod
    def stack_data(X, y, time):
        """[n,P],[n,T] -> rows (x, t) of shape [n*T, P+1] and the flattened targets."""
        n_t = y.shape[-1]
        rows = np.column_stack((np.repeat(X, n_t, axis=0), np.tile(time, len(X))))
        return rows, y.reshape(-1)

    def _fit_scaler(self):
        self.scaler = MinMaxScaler() if self.config.get("SCALING_STRATEGY") == "minmax" else StandardScaler()
        if self._kind == "I":
            tr, self.y_train_scaled = self.stack_data(self.X_train, self.y_train, self.time)
            te, self.y_test_scaled = self.stack_data(self.X_test, self.y_test, self.time)
        else:
            tr, te = self.X_train, self.X_test
            self.y_train_scaled, self.y_test_scaled = self.y_train, self.y_test
        self.X_train_scaled = self.scaler.fit_transform(tr)
        self.X_test_scaled = self.scaler.transform(te)

    def scale_new_data(self, X, y=None):
        if self.time is not None and self._kind == "I":
            rows, y_flat = self.stack_data(X, y, self.time)
            return self.scaler.transform(rows), y_flat
        return self.scaler.transform(X), None


class FixtureProcessor:
    """The same constants without the CSV: ``time`` and a fitted-scaler look-alike (for fixtures)."""

    class _Scaler:
        def __init__(self, mean, scale):
            self.mean_, self.scale_ = np.asarray(mean, np.float64), np.asarray(scale, np.float64)

        def transform(self, X):
            return (np.asarray(X, np.float64) - self.mean_) / self.scale_

    def __init__(self, time, scaler_mean, scaler_scale, config=None):
        self.time = np.asarray(time, np.float64)
        self.scaler = self._Scaler(scaler_mean, scaler_scale)
        self.config = dict(config or {})
