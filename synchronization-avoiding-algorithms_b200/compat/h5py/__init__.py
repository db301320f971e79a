"""h5py for the reference's result files (Data_prepare.py:243-246, Shared_extraction.py:32-40, Online_predictor.py:321-324,
Tools/DNN_tools.py:286-287).

If a real h5py is installed ANYWHERE else on sys.path, this module replaces itself with it at import time, whatever the
order of PYTHONPATH.  Only when none exists (this image ships no HDF5 library at all) the stand-in below is used: the
slice of `h5py.File` those scripts need, reading and writing genuine HDF5 through the package's own `hdf5_lite` — the
files land at exactly the path the caller names and open with h5py / h5dump elsewhere.  (`<path>.npz` archives written by
earlier versions of this stand-in are still read when `<path>` itself does not exist.)
"""
import os as _os
import sys as _sys

import _saa_defer

_real = _saa_defer.real("h5py", __file__)
if _real is not None:
    _sys.modules[__name__] = _real
else:
    import importlib.util as _ilu

    import numpy as _np

    IS_STAND_IN = True
    _src = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))), "hdf5_lite.py")
    _spec = _ilu.spec_from_file_location("saa_hdf5_lite", _src)
    _lite = _ilu.module_from_spec(_spec)
    _sys.modules["saa_hdf5_lite"] = _lite         # registered before execution (dataclass / annotation lookups by module name)
    _spec.loader.exec_module(_lite)

    class File(_lite.File):
        def __init__(self, name, mode="r", **kw):
            legacy = str(name) + ".npz"
            if mode == "r" and not _os.path.isfile(str(name)) and _os.path.isfile(legacy):
                self.filename, self.mode, self._reader, self._addr = str(name), "r", None, {}
                with _np.load(legacy) as z:
                    self._d = {k: z[k] for k in z.files}
                return
            super().__init__(name, mode, **kw)

    Hdf5Error = _lite.Hdf5Error
