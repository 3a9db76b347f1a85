"""Import shim: the package directory the task names is ``mcmc-mems_b200/`` (a hyphen is not
importable), so ``import mcmc_mems_b200`` resolves to that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mcmc-mems_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _fh:
    exec(compile(_fh.read(), __file__, "exec"))
del _os, _fh
