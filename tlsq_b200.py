"""Import shim: the package directory is named ``totalleastsquares.jl_b200`` (with a dot), which the ``import``
statement cannot spell.  ``import tlsq_b200`` loads that directory as a regular package under this name."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "totalleastsquares.jl_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
