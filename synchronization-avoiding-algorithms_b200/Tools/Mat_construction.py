"""Assembly entry points with the reference's names (/root/reference/Tools/Mat_construction.py) on top of the
sparse, vectorised assembly of `saa_b200.assembly` — no dense (3n)^2 intermediate for the stiffness."""
import numpy as np
from scipy.sparse import csr_matrix  # noqa: F401  (re-exported like the reference does)

from saa_b200 import assembly as _asm
from Tools.commons import *            # noqa: F401,F403
from Tools.commons import LumpedCarrier as _LumpedCarrier

_DENSE_LIMIT = 30_000   # DOFs up to which the global helpers hand back true dense arrays like the reference


def Local_assembly_for_stiffness(local_node_list, Cell, Points, deg, n_basis, elas, rank):
    """Per-rank stiffness as scipy CSR (sorted int32 indices, exact zeros dropped) — :122-150."""
    if deg != 1 or n_basis != 4:
        raise NotImplementedError("explicit dynamics uses linear tetrahedra (the reference marks p=2 dynamics TBD)")
    return _asm.local_stiffness_csr(local_node_list, np.asarray(Cell), np.asarray(Points), elas.lmd, elas.mu)


def Global_Assembly_no_bc(deg, Cells, Points, elas, t):
    """(M, K, F) without boundary conditions — :199-231.  F is the (3N,1) load vector.  For small meshes M and
    K are dense arrays; beyond that M is a handle that only supports lumping_to_vec() and K is CSR."""
    if deg != 1:
        raise NotImplementedError("p = 1 only")
    Points, Cells = np.asarray(Points), np.asarray(Cells)
    n = 3 * len(Points)
    f = elas.f(None, t)
    lM, F = _asm.lumped_mass_and_load(Points, Cells, elas.rho, -float(f[2, 0]))
    K = _asm.local_stiffness_csr(np.arange(len(Points)), Cells, Points, elas.lmd, elas.mu)
    if n <= _DENSE_LIMIT:
        M = _asm.consistent_mass_csr(Points, Cells, elas.rho).toarray().view(_LumpedCarrier)
        M.row_sums = lM
        return M, K.toarray(), F
    M = np.zeros((0, 0)).view(_LumpedCarrier)
    M.row_sums = lM
    return M, K, F


def Global_Assembly(deg, Cells, Points, Dirichlet, elas, t, Facets=None, Neumann=None, steady=False):
    """(M, K, F) with the rows / columns of clamped DOFs left empty — :155-196."""
    if deg != 1:
        raise NotImplementedError("p = 1 only")
    Points, Cells = np.asarray(Points), np.asarray(Cells)
    n = 3 * len(Points)
    if n > _DENSE_LIMIT:
        raise MemoryError("Global_Assembly returns dense (3N)^2 arrays; it is a set-up helper for small meshes")
    f = elas.f(None, t)
    _, F = _asm.lumped_mass_and_load(Points, Cells, elas.rho, -float(f[2, 0]))
    K = _asm.local_stiffness_csr(np.arange(len(Points)), Cells, Points, elas.lmd, elas.mu).toarray()
    M = _asm.consistent_mass_csr(Points, Cells, elas.rho).toarray()
    D = np.asarray(Dirichlet, dtype=np.int64)
    for A in (M, K):
        A[D, :] = 0.0
        A[:, D] = 0.0
    F = F.copy()
    F[D] = 0.0
    return M, K, F
