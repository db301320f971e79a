"""Make the product package importable as `saa_b200` (its directory name is not a valid identifier)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

if "saa_b200" not in _sys.modules:
    _pkg = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    _spec = _ilu.spec_from_file_location("saa_b200", _os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
    _mod = _ilu.module_from_spec(_spec)
    _sys.modules["saa_b200"] = _mod
    _spec.loader.exec_module(_mod)
