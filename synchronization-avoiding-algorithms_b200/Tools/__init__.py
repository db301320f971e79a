"""Drop-in `Tools` package: the reference's call surface for the explicit time-step path.

Put the directory that contains this package (`synchronization-avoiding-algorithms_b200/`) in front of
`sys.path` and the reference drivers' `from Tools.commons import *`, `from Tools.Distributed_tools import *`,
`from Tools.Dynamic_solver import *` ... (/root/reference/Data_prepare.py:1-4, Online_predictor.py:2-6) resolve
to these modules: same function names, argument meaning and return shapes, with the per-step work done by
the sm_100a kernels behind include/saa_fem.h.  See INTEGRATION.md.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

if "saa_b200" not in _sys.modules:          # make the package importable under its alias (see saa_b200.py)
    _pkg = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    _spec = _ilu.spec_from_file_location("saa_b200", _os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
    _mod = _ilu.module_from_spec(_spec)
    _sys.modules["saa_b200"] = _mod
    _spec.loader.exec_module(_mod)
