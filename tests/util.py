"""Shared helpers of the test-suite: golden fixture loading and oracle construction."""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*_P*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    P = int(g["size"])
    ranks = []
    for q in range(P):
        ranks.append(dict(K_indptr=g[f"r{q}_K_indptr"], K_indices=g[f"r{q}_K_indices"], K_data=g[f"r{q}_K_data"],
                          F=g[f"r{q}_F"], lM=g[f"r{q}_lM"], dirichlet=g[f"r{q}_dirichlet"], nodes=g[f"r{q}_nodes"],
                          ele=g[f"r{q}_ele"], shared=g[f"r{q}_shared"], loc_dof_shared=g[f"r{q}_loc_dof_shared"]))
    g["ranks"] = ranks
    g["P"] = P
    return g


def oracle_module():
    """oracle/fem_oracle.py — the CPU checker (tests may import it; the product never does)."""
    p = os.path.join(ROOT, "oracle")
    if p not in sys.path:
        sys.path.insert(0, p)
    import fem_oracle
    return fem_oracle


def make_oracle(g):
    """OracleProblem (oracle/fem_oracle.c) for a golden case."""
    return oracle_module().OracleProblem(len(g["points"]), g["ranks"], g["dt"], float(g["alpha"]))


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def read_result(path, name="Displacement"):
    """Dataset `name` of a result file the drivers wrote with h5py.File(path, 'w') (Data_prepare.py:243-246): genuine
    HDF5 at `path` (real h5py, or the package's hdf5_lite behind the compat stand-in); `<path>.npz` archives of the
    earlier stand-in are accepted too."""
    if os.path.isfile(str(path)):
        import saa_b200  # noqa: F401
        from saa_b200 import hdf5_lite
        return hdf5_lite.read_file(str(path))[name]
    return np.load(str(path) + ".npz")[name]


def write_result(path, array, name="Displacement"):
    """A result file as the drivers write it (genuine HDF5, contiguous float64)."""
    import saa_b200  # noqa: F401
    from saa_b200 import hdf5_lite
    os.makedirs(os.path.dirname(str(path)) or ".", exist_ok=True)
    hdf5_lite.write_file(str(path), {name: np.asarray(array)})
