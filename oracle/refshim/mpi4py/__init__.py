"""Single-process stand-in for mpi4py (oracle harness only)."""
from . import MPI  # noqa: F401
