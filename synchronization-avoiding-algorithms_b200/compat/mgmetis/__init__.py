"""Stand-in for mgmetis: see compat/README.md."""
import _saa_defer

_real = _saa_defer.real("mgmetis", __file__)
if _real is not None:
    import sys as _sys
    _sys.modules[__name__] = _real
