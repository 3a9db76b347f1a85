"""CPU oracle for the batched surrogate-MH hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy, the algorithm the reference runs for one
Metropolis-Hastings step: the forward-model closure of
``src/InverseProblems/utils.py:22-42``, the Keras ``Dense``/tanh stack of
``src/SurrogateModeling/model.py:56-59`` and the CUQIpy 0.8.0 sampler /
distribution arithmetic the notebooks call
(``tests/Xaccelerometer_geometric/identification.ipynb:136-169,292-308``).

Third-party arithmetic restated here because it is NOT vendored in the reference
tree: ``cuqipy==0.8.0`` (README.md:36) and ``tensorflow==2.7.1`` (README.md:38).

Pinning: the oracle is checked against every known-answer artefact the reference
holds for this path: TF float32 outputs printed in
``tests/Xaccelerometer_quality_factor/surrogate_fft.ipynb`` cell 11, the
``X_test[10]`` pins, the printed acceptance rates / adapted scales and the golden
posterior ``samples/sample_10.npy``.  Delayed acceptance and component-wise MH
have no reference implementation or artefact: PARITY UNPINNED for those two
(they are defined from Christen & Fox 2005 and CUQIpy's ``CWMH`` semantics).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
(``mcmc-mems_b200/``) never does.
"""
from .surrogate import mlp_forward, make_forward_model, batched_forward  # noqa: F401
from .mh import (  # noqa: F401
    gaussian_loglike, uniform_logpdf, make_posterior_logd, decide,
    mh_sample_adapt, mh_sample, cwmh_sample, da_sample,
)
from .philox import philox4x32_10  # noqa: F401
