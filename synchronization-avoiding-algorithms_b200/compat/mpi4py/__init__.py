"""Stand-in for mpi4py (used only when the real package is not installed): see compat/README.md."""
import _saa_defer

_real = _saa_defer.real("mpi4py", __file__)
if _real is not None:
    import sys as _sys
    _sys.modules[__name__] = _real
else:
    from . import MPI  # noqa: F401
