"""`from mpi4py import MPI` -> COMM_WORLD over torch.distributed (saa_b200.comm.TorchDistComm) or serial."""
import _saa_bootstrap  # noqa: F401,E402

_saa_stub = True
from saa_b200 import comm as _comm  # noqa: E402

COMM_WORLD = _comm.world()
DOUBLE = "d"
INT64_T = "q"
