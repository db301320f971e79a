"""A compat stand-in steps aside for the real package: `real(name, __file__)` looks `name` up on sys.path WITHOUT this
directory and, when found, imports it, installs it as sys.modules[name] and returns it (None otherwise).  Stand-ins call
this first thing, so a real mpi4py / meshio / mgmetis / matplotlib / h5py always wins — whatever the order of PYTHONPATH."""
import importlib.machinery as _mach
import importlib.util as _ilu
import os as _os
import sys as _sys

_HERE = _os.path.dirname(_os.path.abspath(__file__))


def real(name, stub_file):
    paths = [p for p in _sys.path if _os.path.abspath(p or ".") != _HERE]
    try:
        spec = _mach.PathFinder.find_spec(name, paths)
    except Exception:
        spec = None
    if spec is None or spec.origin is None or _os.path.abspath(spec.origin) == _os.path.abspath(stub_file):
        return None
    mod = _ilu.module_from_spec(spec)
    saved = _sys.modules.get(name)
    _sys.modules[name] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        if saved is not None:
            _sys.modules[name] = saved
        else:
            _sys.modules.pop(name, None)
        return None
    return mod
