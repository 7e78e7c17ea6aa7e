"""Import shim: the package directory is named `neuralgraphpde.jl_b200` (not a valid Python identifier), so
`import ngpde` loads it under the name `ngpde`."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "neuralgraphpde.jl_b200")
_spec = importlib.util.spec_from_file_location("ngpde", os.path.join(_pkg_dir, "__init__.py"),
                                               submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ngpde"] = _mod
_spec.loader.exec_module(_mod)
