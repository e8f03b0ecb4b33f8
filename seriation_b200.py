"""Import shim: loads the package directory ``seriation-in-paleontological-data-using-mcmc_b200/``
(not a valid Python identifier) under the module name ``seriation_b200``."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "seriation-in-paleontological-data-using-mcmc_b200")
_NAME = "seriation_b200_pkg"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

_pkg = sys.modules[_NAME]
api = _pkg.api
globals().update({k: getattr(api, k) for k in dir(api) if not k.startswith("_")})
PKG_DIR = _PKG_DIR
