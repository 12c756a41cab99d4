"""Importable alias for the package directory `itensornumericalanalysis.jl_b200/` (a dot in a
directory name cannot be imported directly).  `import itna_b200` yields that package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "itensornumericalanalysis.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "itna_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["itna_b200"] = _mod
_spec.loader.exec_module(_mod)
