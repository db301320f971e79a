"""Steady solve used by the drivers' set-up (/root/reference/Tools/Steady_solvers.py:13-22) — sparse."""
import numpy as np
from scipy.sparse import csr_matrix, identity
from scipy.sparse.linalg import spsolve

from saa_b200 import assembly as _asm
from Tools.Mat_construction import *   # noqa: F401,F403
from Tools.commons import *            # noqa: F401,F403

pi = np.pi


def Steady_Elasticity_solver(p, Cells, Points, Dirichlet, elas, t=None, Facets=None, Neumann=None):
    """Solve K d = F with the clamped DOFs eliminated (identity rows) — returns d (3N,1)."""
    Points, Cells = np.asarray(Points), np.asarray(Cells)
    n = 3 * len(Points)
    f = elas.f(None, t)
    _, F = _asm.lumped_mass_and_load(Points, Cells, elas.rho, -float(f[2, 0]))
    K = _asm.local_stiffness_csr(np.arange(len(Points)), Cells, Points, elas.lmd, elas.mu).tocsr().astype(np.float64)
    keep = np.ones(n)
    keep[np.asarray(Dirichlet, dtype=np.int64)] = 0.0
    S = csr_matrix((keep, (np.arange(n), np.arange(n))), shape=(n, n))
    A = S @ K @ S + (identity(n, format="csr") - S)
    return spsolve(A.tocsc(), (F[:, 0] * keep)).reshape(n, 1)
