"""Stand-in for h5py backed by .npz (same file name + '.npz'): see compat/README.md."""
import os

import numpy as np


class File:
    def __init__(self, name, mode="r"):
        self.name, self.mode, self._d = str(name), mode, {}
        if mode.startswith("r"):
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
            os.makedirs(os.path.dirname(self.name) or ".", exist_ok=True)
            np.savez_compressed(self.name + ".npz", **self._d)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
