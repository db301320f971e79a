"""Sparse (never dense) assembly of the per-rank stiffness, lumped mass and load vector.

The reference accumulates into dense (3n)x(3n) arrays and converts with csr_matrix(K)
(/root/reference/Tools/Mat_construction.py:122-150, 199-231), which is O(n^2) memory.  Here the
same numbers are produced without any dense matrix:

* element matrices are evaluated for all elements at once with the *same numpy calls* the
  reference makes per element (np.linalg.det / inv on the 3x3 Jacobian, `Bi.T @ D @ Bj * detJ * w`,
  Mat_construction.py:90-117) applied to stacks, so each 3x3 block gets the value the reference
  computes;
* contributions to one matrix entry are added in ascending local element order starting from
  0.0, which is the order of the `K[P,Q] += Local_Ke[p,q]` loop (Mat_construction.py:125-148);
* exact zeros are dropped and columns are ascending within a row, as csr_matrix(dense) does
  (Mat_construction.py:150); indices are int32 like scipy's.

`tests/test_maps_assembly.py` checks bit-equality with the reference's own output (in the authoring
container) and ulp-level agreement with the golden fixtures elsewhere.
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import csr_matrix

# 4-point rule of Tools/Qudrature.py:7-12: every weight is 0.25/6; for linear tets the integrand
# of K is constant, so the four quadrature terms are four equal addends.
_W_QUAD = np.array([0.25 / 6, 0.25 / 6, 0.25 / 6, 0.25 / 6])
_N_QUAD_POINTS = np.array([[0.5854101966249685, 0.1381966011250105, 0.1381966011250105],
                           [0.1381966011250105, 0.5854101966249685, 0.1381966011250105],
                           [0.1381966011250105, 0.1381966011250105, 0.5854101966249685],
                           [0.1381966011250105, 0.1381966011250105, 0.1381966011250105]])
# Tools/Shape_function_Deriv.py:36 (p = 1)
_D_XI = np.array([[-1.0, -1.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])


def elasticity_D(lmd, mu):
    """Tools/commons.py:25-31."""
    return np.array([[lmd + 2.0 * mu, lmd, lmd, 0.0, 0.0, 0.0],
                     [lmd, lmd + 2.0 * mu, lmd, 0.0, 0.0, 0.0],
                     [lmd, lmd, lmd + 2.0 * mu, 0.0, 0.0, 0.0],
                     [0.0, 0.0, 0.0, mu, 0.0, 0.0],
                     [0.0, 0.0, 0.0, 0.0, mu, 0.0],
                     [0.0, 0.0, 0.0, 0.0, 0.0, mu]])


def _jacobians(P):
    """Shape_function_Deriv.py:60-67 on a stack: J[e,i,j] = dot(D_xi[:,j], P[e,:,i]).

    With D_xi = [[-1,-1,-1],[1,0,0],[0,1,0],[0,0,1]] the dot has two exact-zero terms and two
    terms that are exact products by +-1, so its value is fl(P[e,j+1,i] - P[e,0,i]) in any order.
    """
    return np.transpose(P[:, 1:4, :] - P[:, 0:1, :], (0, 2, 1)).copy()


def _B_stack(N_xyz):
    """Mat_construction.py:99-104 for all elements and all 4 local nodes: (nE,4,6,3)."""
    nE = N_xyz.shape[0]
    B = np.zeros((nE, 4, 6, 3))
    gx, gy, gz = N_xyz[:, :, 0], N_xyz[:, :, 1], N_xyz[:, :, 2]
    B[:, :, 0, 0] = gx
    B[:, :, 1, 1] = gy
    B[:, :, 2, 2] = gz
    B[:, :, 3, 1] = gz
    B[:, :, 3, 2] = gy
    B[:, :, 4, 0] = gz
    B[:, :, 4, 2] = gx
    B[:, :, 5, 0] = gy
    B[:, :, 5, 1] = gx
    return B


def element_stiffness_blocks(points, cells, lmd, mu):
    """3x3 blocks Ke[e,a,b,A,B] == Local_K_coronary(...)[3a+A,3b+B] (Mat_construction.py:79-119).

    Returned shape (nE,4,4,3,3).
    """
    P = points[np.asarray(cells)[:, :4]]                  # (nE,4,3) like Pt of :135
    Jac = _jacobians(P)
    detJ = np.linalg.det(Jac)                             # :93
    invJ = np.linalg.inv(Jac)                             # :94
    N_xyz = _D_XI @ invJ                                  # :96  (4x3)@(3x3) per element
    B = _B_stack(N_xyz)
    D = elasticity_D(lmd, mu)
    nE = P.shape[0]
    Ke = np.zeros((nE, 4, 4, 3, 3))
    w = _W_QUAD[0]
    for i in range(4):
        BiT = np.transpose(B[:, i], (0, 2, 1))            # view, like np.transpose(Bi) at :112
        BiTD = BiT @ D
        for j in range(4):
            k_loc = BiTD @ B[:, j] * detJ[:, None, None] * w      # :112, left-to-right
            # four identical quadrature addends, K starts at 0.0 (:82, :117)
            acc = np.zeros_like(k_loc)
            for _ in range(4):
                acc = acc + k_loc
            Ke[:, i, j] = acc
    return Ke


def _grouped_sequential_sum(keys_sorted_group_start, vals_sorted):
    """Sum runs of `vals_sorted` (runs start where keys_sorted_group_start is True) strictly left to
    right, each run starting from 0.0: (((0+v0)+v1)+v2)...  vals may carry trailing dims."""
    n = vals_sorted.shape[0]
    starts = np.nonzero(keys_sorted_group_start)[0]
    lens = np.diff(np.append(starts, n))
    out = np.zeros((starts.size,) + vals_sorted.shape[1:])
    k = 0
    live = np.arange(starts.size)
    while live.size:
        out[live] = out[live] + vals_sorted[starts[live] + k]
        k += 1
        live = live[lens[live] > k]
    return out


def local_stiffness_csr(local_node_list, cells_local, points, lmd, mu, chunk=None):
    """Sparse restatement of Local_assembly_for_stiffness (Mat_construction.py:122-150).

    local_node_list: global node ids in the rank's local order (row/col 3k+A belongs to entry k);
    cells_local:     (nE_loc,4) global node ids of the rank's elements in Local_ele_list order.
    Returns scipy csr_matrix (3n x 3n), float64 data, int32 indices/indptr, sorted columns, no
    stored zeros.
    """
    L = np.asarray(local_node_list, dtype=np.int64)
    cells_local = np.asarray(cells_local, dtype=np.int64)
    n = L.size
    nE = cells_local.shape[0]
    # global id -> local position (local_mat_node, Distributed_tools.py:66-73)
    order = np.argsort(L, kind="stable")
    loc = order[np.searchsorted(L[order], cells_local)]   # (nE,4) local node positions
    Ke = element_stiffness_blocks(points, cells_local, lmd, mu)      # (nE,4,4,3,3)

    # node-block COO: key = row_node * n + col_node, contributions kept in element order
    rn = np.repeat(loc[:, :, None], 4, axis=2).reshape(-1)
    cn = np.repeat(loc[:, None, :], 4, axis=1).reshape(-1)
    key = rn * np.int64(n) + cn
    perm = np.argsort(key, kind="stable")                 # stable: ascending element order inside a key
    key_s = key[perm]
    start = np.empty(key_s.size, dtype=bool)
    start[0] = True
    start[1:] = key_s[1:] != key_s[:-1]
    blocks = _grouped_sequential_sum(start, Ke.reshape(-1, 3, 3)[perm])   # (nblk,3,3)
    bkey = key_s[start]
    brow, bcol = bkey // n, bkey % n                      # sorted by (row node, col node)

    # expand to scalar entries in CSR order: row 3*brow+A, cols ascending = (bcol, B)
    nblk = bkey.size
    blk_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(blk_ptr, brow + 1, 1)
    np.cumsum(blk_ptr, out=blk_ptr)
    indptr_full = np.zeros(3 * n + 1, dtype=np.int64)
    nb_per_node = np.diff(blk_ptr)
    indptr_full[1:] = np.cumsum(np.repeat(nb_per_node * 3, 3))
    # position of block j inside its node-row
    within = np.arange(nblk, dtype=np.int64) - blk_ptr[brow]
    data = np.empty(indptr_full[-1])
    indices = np.empty(indptr_full[-1], dtype=np.int64)
    for A in range(3):
        base = indptr_full[3 * brow + A] + 3 * within
        for B_ in range(3):
            data[base + B_] = blocks[:, A, B_]
            indices[base + B_] = 3 * bcol + B_
    # csr_matrix(dense) keeps only entries != 0 (Mat_construction.py:150)
    keep = data != 0
    row_of = np.repeat(np.arange(3 * n, dtype=np.int64), np.diff(indptr_full))
    indptr = np.zeros(3 * n + 1, dtype=np.int64)
    np.add.at(indptr, row_of[keep] + 1, 1)
    np.cumsum(indptr, out=indptr)
    idx_dtype = np.int32 if max(3 * n, int(indptr[-1])) < 2 ** 31 else np.int64
    K = csr_matrix((data[keep], indices[keep].astype(idx_dtype), indptr.astype(idx_dtype)), shape=(3 * n, 3 * n))
    K.has_sorted_indices = True
    return K


def _pairwise_sum_sparse(lo, n, pos, val):
    """Value of numpy's pairwise float64 summation (np.sum over a contiguous row, as used by
    lumping_to_vec, commons.py:103-107) of a length-n row that is zero except val[k] at pos[k]
    (ascending), without materialising the row.  Mirrors numpy's scheme: n < 8 sequential; n <= 128
    eight strided accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail;
    otherwise split at n//2 rounded down to a multiple of 8.  Adding exact zeros never rounds, so
    only the non-zeros' places in that tree matter."""
    m = len(pos)
    if m == 0:
        return 0.0
    if m == 1:
        return val[0]
    if n < 8:
        r = -0.0
        for v in val:
            r = r + v
        return r
    if n <= 128:
        main = n - (n % 8)
        acc = [0.0] * 8
        tail = []
        for p_, v in zip(pos, val):
            i = p_ - lo
            if i < main:
                acc[i % 8] = acc[i % 8] + v
            else:
                tail.append(v)
        res = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]))
        for v in tail:
            res = res + v
        return res
    n2 = n // 2
    n2 -= n2 % 8
    k = int(np.searchsorted(pos, lo + n2))
    return _pairwise_sum_sparse(lo, n2, pos[:k], val[:k]) + _pairwise_sum_sparse(lo + n2, n - n2, pos[k:], val[k:])


def lumped_mass_and_load(points, cells, rho, fz, exact_rowsum=None):
    """Lumped mass vector and un-ramped load vector over the WHOLE mesh (global DOF numbering).

    Restates Global_Assembly_no_bc + lumping_to_vec (Mat_construction.py:199-231, commons.py:103-107,
    Data_prepare.py:175-176) sparsely.  Element values follow Local_MKF (Mat_construction.py:36-73):
    m_loc = N_i*rho*N_j*detJ*w summed over the 4 quadrature points, F_e = N_i*f_C*detJ*w.  Element
    contributions are added in ascending element order.  The reference then row-sums the dense
    global mass with numpy's pairwise np.sum over all 3N columns.  With exact_rowsum=True (default
    for 3N <= 300k) that pairwise tree is evaluated on the stored entries (`_pairwise_sum_sparse`), so
    l_M is bit-identical to the reference; otherwise the stored entries are added in ascending column
    order, which may differ in the last bit (|rel| <= 3e-16) — irrelevant beyond the sizes the
    reference itself can run (dense (3N)^2 arrays).
    Returns (l_M (3N,1), F_pre (3N,1)).
    """
    cells = np.asarray(cells, dtype=np.int64)
    N = points.shape[0]
    P = points[cells[:, :4]]
    detJ = np.linalg.det(_jacobians(P))
    xi = _N_QUAD_POINTS
    shp = np.stack([1. - xi[:, 0] - xi[:, 1] - xi[:, 2], xi[:, 0], xi[:, 1], xi[:, 2]], axis=1)  # (4qp,4) Shape_Function
    nE = cells.shape[0]
    # consistent mass block (scalar per node pair), quadrature sum in order, starting at 0.0
    Me = np.zeros((nE, 4, 4))
    Fe = np.zeros((nE, 4, 3))
    f_loc = np.array([0.0, -fz, -fz])                     # commons.py:35-38 (R == False)
    for q in range(4):
        for i in range(4):
            for j in range(4):
                Me[:, i, j] = Me[:, i, j] + shp[q, i] * rho * shp[q, j] * detJ * _W_QUAD[q]   # :62,:68
            for C in range(3):
                Fe[:, i, C] = Fe[:, i, C] + shp[q, i] * f_loc[C] * detJ * _W_QUAD[q]           # :73
    # F[P] += Fe[p] in element order, loop A (dir) outer, a (node) inner — each (node,dir) gets one
    # addend per element, so only the element order matters
    node = cells.reshape(-1)
    perm = np.argsort(node, kind="stable")
    ns = node[perm]
    start = np.empty(ns.size, dtype=bool)
    start[0] = True
    start[1:] = ns[1:] != ns[:-1]
    Fsum = _grouped_sequential_sum(start, Fe.reshape(-1, 3)[perm])
    F = np.zeros((N, 3))
    F[ns[start]] = Fsum
    # mass: M[P,Q] += Me[p,q] in element order per (node pair), then row sum
    rn = np.repeat(cells[:, :, None], 4, axis=2).reshape(-1)
    cn = np.repeat(cells[:, None, :], 4, axis=1).reshape(-1)
    key = rn * np.int64(N) + cn
    perm = np.argsort(key, kind="stable")
    ks = key[perm]
    start = np.empty(ks.size, dtype=bool)
    start[0] = True
    start[1:] = ks[1:] != ks[:-1]
    Mpair = _grouped_sequential_sum(start, Me.reshape(-1)[perm])
    prow, pcol = ks[start] // N, ks[start] % N
    rstart = np.empty(prow.size, dtype=bool)
    rstart[0] = True
    rstart[1:] = prow[1:] != prow[:-1]
    if exact_rowsum is None:
        exact_rowsum = 3 * N <= 300_000
    if exact_rowsum:
        # row 3a+A of the dense global mass holds M(a,b) at column 3b+A: the positions, hence the
        # pairwise tree, depend on the component A
        l_M = np.zeros((3 * N, 1))
        bounds = np.append(np.nonzero(rstart)[0], prow.size)
        for g in range(bounds.size - 1):
            a = int(prow[bounds[g]])
            cols = pcol[bounds[g]:bounds[g + 1]]
            vals = Mpair[bounds[g]:bounds[g + 1]].tolist()
            for A in range(3):
                l_M[3 * a + A, 0] = _pairwise_sum_sparse(0, 3 * N, 3 * cols + A, vals)
    else:
        Mrow = _grouped_sequential_sum(rstart, Mpair)
        lM = np.zeros(N)
        lM[prow[rstart]] = Mrow
        l_M = np.repeat(lM, 3).reshape(3 * N, 1)
    return l_M, F.reshape(3 * N, 1)


def consistent_mass_csr(points, cells, rho):
    """Consistent mass matrix of the whole mesh as CSR (the M of Global_Assembly*, Mat_construction.py:62,68,
    199-231): scalar node-pair values Me[a,b] = sum_q N_a rho N_b detJ w on the three diagonal components,
    element contributions added in ascending element order."""
    cells = np.asarray(cells, dtype=np.int64)
    N = points.shape[0]
    P = points[cells[:, :4]]
    detJ = np.linalg.det(_jacobians(P))
    xi = _N_QUAD_POINTS
    shp = np.stack([1. - xi[:, 0] - xi[:, 1] - xi[:, 2], xi[:, 0], xi[:, 1], xi[:, 2]], axis=1)
    Me = np.zeros((cells.shape[0], 4, 4))
    for q in range(4):
        for i in range(4):
            for j in range(4):
                Me[:, i, j] = Me[:, i, j] + shp[q, i] * rho * shp[q, j] * detJ * _W_QUAD[q]
    rn = np.repeat(cells[:, :, None], 4, axis=2).reshape(-1)
    cn = np.repeat(cells[:, None, :], 4, axis=1).reshape(-1)
    key = rn * np.int64(N) + cn
    perm = np.argsort(key, kind="stable")
    ks = key[perm]
    start = np.empty(ks.size, dtype=bool)
    start[0] = True
    start[1:] = ks[1:] != ks[:-1]
    Mpair = _grouped_sequential_sum(start, Me.reshape(-1)[perm])
    prow, pcol = ks[start] // N, ks[start] % N
    rows = (3 * prow[:, None] + np.arange(3)[None, :]).ravel()
    cols = (3 * pcol[:, None] + np.arange(3)[None, :]).ravel()
    return csr_matrix((np.repeat(Mpair, 3), (rows, cols)), shape=(3 * N, 3 * N))
