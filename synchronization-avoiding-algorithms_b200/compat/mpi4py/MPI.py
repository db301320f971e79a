"""`from mpi4py import MPI` -> COMM_WORLD over torch.distributed (saa_b200.comm.TorchDistComm) or serial."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_saa_stub = True
if "saa_b200" not in _sys.modules:
    _pkg = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
    _spec = _ilu.spec_from_file_location("saa_b200", _os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
    _mod = _ilu.module_from_spec(_spec)
    _sys.modules["saa_b200"] = _mod
    _spec.loader.exec_module(_mod)
from saa_b200 import comm as _comm  # noqa: E402

COMM_WORLD = _comm.world()
DOUBLE = "d"
INT64_T = "q"
