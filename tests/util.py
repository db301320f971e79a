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
