"""Import shim: the package directory `radialbasisfinitedifferences.jl_b200/` (name fixed by the project layout)
contains a dot, so it is loaded by path and registered as module `rbffd_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "radialbasisfinitedifferences.jl_b200")
_spec = importlib.util.spec_from_file_location("rbffd_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["rbffd_b200"] = _mod
_spec.loader.exec_module(_mod)
