"""h5py for the reference's result files (Data_prepare.py:243-246, Shared_extraction.py:38-40, Online_predictor.py:321-324).

If a real h5py is installed ANYWHERE else on sys.path, this module replaces itself with it at import time — genuine
HDF5 files are written, whatever the order of PYTHONPATH.  Only when none exists (this image ships no HDF5 library at
all) the minimal stand-in below is used: `File(path).create_dataset(name, data=...)` / `File(path)[name]` backed by ONE
`<path>.npz` per file (path "Results/Dynamics/Local-rank-0.hdf5" -> "Results/Dynamics/Local-rank-0.hdf5.npz", dataset
names = npz keys, `np.load(...)["Displacement"]` reads it).  The substitution is announced once per process on stderr:
such files cannot be opened by tools that expect HDF5.
"""
import os as _os
import sys as _sys

import _saa_defer

_real = _saa_defer.real("h5py", __file__)
if _real is not None:
    _sys.modules[__name__] = _real
else:
    import numpy as np

    IS_STAND_IN = True
    _warned = False

    def _warn(path):
        global _warned
        if not _warned:
            _warned = True
            print(f"[saa_b200.compat.h5py] h5py is not installed: result files are written as NumPy archives, "
                  f"'{path}' -> '{path}.npz' (same dataset names).  Install h5py to get genuine HDF5.", file=_sys.stderr)

    class File:
        def __init__(self, name, mode="r"):
            self.name, self.mode, self._d = str(name), mode, {}
            if mode.startswith("r"):
                if _os.path.isfile(self.name) and not _os.path.isfile(self.name + ".npz"):
                    raise OSError(f"{self.name} exists but h5py is not installed (this stand-in reads only {self.name}.npz)")
                with np.load(self.name + ".npz") as z:
                    self._d = {k: z[k] for k in z.files}

        def create_dataset(self, name, data=None, compression=None, **_):
            self._d[name] = np.asarray(data)
            return self._d[name]

        def __getitem__(self, k):
            return self._d[k]

        def keys(self):
            return self._d.keys()

        def close(self):
            if not self.mode.startswith("r"):
                _warn(self.name)
                _os.makedirs(_os.path.dirname(self.name) or ".", exist_ok=True)
                np.savez_compressed(self.name + ".npz", **self._d)

        def __enter__(self):
            return self

        def __exit__(self, *a):
            self.close()
