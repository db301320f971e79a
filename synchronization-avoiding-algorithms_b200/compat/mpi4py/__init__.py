"""Stand-in for mpi4py (used only when the real package is not installed): see compat/README.md."""
from . import MPI  # noqa: F401
