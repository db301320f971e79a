"""Stand-in for matplotlib (used only when the real package is not installed): plotting calls are accepted and
ignored, `savefig` writes nothing.  See compat/README.md."""
import _saa_defer

_real = _saa_defer.real("matplotlib", __file__)
if _real is not None:
    import sys as _sys
    _sys.modules[__name__] = _real
