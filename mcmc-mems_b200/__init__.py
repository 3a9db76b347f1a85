"""mcmc-mems_b200: B200-native batched surrogate Metropolis-Hastings, drop-in for the hot path of
filippozacchei/MCMC-MEMS.  Host side is Python mirroring the reference's own
interfaces; all arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/mems_b200.h`` (``libmems_b200.so``, loaded with ctypes).  No CPU fallback.
"""
from ._lib import MemsError, LIB_PATH, PATH_FP32, PATH_TC  # noqa: F401
from .engine import MHEngine, SurrogateSpec, mlp_forward  # noqa: F401
from .model import NN_Model  # noqa: F401
from .preprocessing import preprocessing, FixtureProcessor  # noqa: F401
from .utils import (create_forward_model_function, least_squares_optimization,  # noqa: F401
                    objective_function, surrogate_spec, batched_least_squares, sensitivity_push_forward)
from .cuqi_compat import (Gaussian, Uniform, JointDistribution, Posterior, Model, MH, CWMH,  # noqa: F401
                          DelayedAcceptanceMH, Samples, BatchedSamples, Continuous1D, Discrete)
from . import diagnostics, diagnostics_device, distributed, fixtures, h5lite  # noqa: F401

__all__ = [
    "MHEngine", "SurrogateSpec", "NN_Model", "preprocessing", "FixtureProcessor",
    "create_forward_model_function", "least_squares_optimization", "batched_least_squares", "sensitivity_push_forward", "Gaussian", "Uniform",
    "JointDistribution", "Posterior", "Model", "MH", "CWMH", "DelayedAcceptanceMH", "Samples", "BatchedSamples", "Continuous1D",
    "Discrete", "diagnostics", "diagnostics_device", "h5lite", "mlp_forward", "MemsError", "PATH_FP32", "PATH_TC",
]
