"""Surrogate wrapper with the reference's ``NN_Model`` surface
(src/SurrogateModeling/model.py:16-34,142-151): ``load_model(path)``, ``predict(X)`` and a
callable ``.model(X)`` whose result has ``.numpy()`` (that is how the forward closure uses it,
src/InverseProblems/utils.py:39).  Weights come from the Keras HDF5 file through ``h5lite``;
inference runs on the GPU through the C ABI (``mems_mlp_forward``).  Training, building and
saving models stay with the reference's offline TensorFlow tooling (out of scope).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from . import h5lite


class _DeviceResult:
    """What ``nn_model.model(X)`` returns: a device tensor exposing ``.numpy()`` like a TF EagerTensor."""

    def __init__(self, t):
        self.tensor = t

    def numpy(self) -> np.ndarray:
        return self.tensor.detach().cpu().numpy()

    @property
    def shape(self):
        return tuple(self.tensor.shape)


class DenseStack:
    """A loaded Keras ``Sequential`` of ``Dense`` layers, callable like the Keras model object."""

    def __init__(self, layers: List[h5lite.DenseLayer], device: int = 0):
        self.layer_objs = layers
        self.device = device

    @property
    def layers(self) -> List[Tuple[np.ndarray, np.ndarray, str]]:
        return [(l.kernel, l.bias, l.activation) for l in self.layer_objs]

    @property
    def input_shape(self):
        return (None, self.layer_objs[0].kernel.shape[0])

    @property
    def output_shape(self):
        return (None, self.layer_objs[-1].kernel.shape[1])

    def get_weights(self) -> List[np.ndarray]:
        out = []
        for l in self.layer_objs:
            out += [l.kernel, l.bias]
        return out

    def __call__(self, X) -> _DeviceResult:
        from .engine import mlp_forward
        X = np.asarray(X)
        if X.ndim != 2 or X.shape[1] != self.input_shape[1]:
            raise ValueError(f"expected input of shape [n,{self.input_shape[1]}], got {X.shape}")
        return _DeviceResult(mlp_forward(self.layers, X.astype(np.float32), self.device))

    def predict(self, X, verbose: int = 0) -> np.ndarray:
        return self(X).numpy()


class NN_Model:
    """Drop-in for the reference class of the same name (inference surface only)."""

    def __init__(self, device: int = 0):
        self.model: Optional[DenseStack] = None
        self.history = None
        self.device = device

    def load_model(self, model_path: str) -> None:
        self.model = DenseStack(h5lite.load_keras_dense_stack(model_path), self.device)

    def set_layers(self, layers) -> None:
        """Install weights directly (``[(kernel, bias, activation), ...]``), e.g. from a fixture."""
        objs = [h5lite.DenseLayer(f"dense_{i}", np.asarray(k, np.float32), np.asarray(b, np.float32), a)
                for i, (k, b, a) in enumerate(layers)]
        self.model = DenseStack(objs, self.device)

    def predict(self, X: np.ndarray) -> np.ndarray:
        if self.model is None:
            raise ValueError("no model loaded: call load_model(path) first")
        return self.model.predict(X, verbose=0)

    def build_model(self, *a, **k):
        raise NotImplementedError("surrogate training is offline TensorFlow tooling in the reference; out of scope here")

    train_model = save_model = plot_training_history = build_model
