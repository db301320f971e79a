"""Meshes for the explicit FE path: legacy-VTK reader and the structured cantilever generator.

* `read_vtk` replaces the `meshio.read` call of /root/reference/Data_prepare.py:58-61 for the
  legacy ASCII files gmsh writes (Mesh_info/beam_coarse.vtk): points (N,3) f64, tetra (nE,4) and
  boundary triangle (nF,3) connectivity as int64.
* `structured_beam(m)` realises Mesh_info/beam_US.geo (a 25 x 1 x 1 box, :1-30) as a structured
  tetrahedral mesh (gmsh is not available offline): 25m x m x m hexahedra, each cut into 6 Kuhn
  tetrahedra with positive Jacobian (the reference integrates with the *signed* det J,
  Tools/Mat_construction.py:93,112), plus the boundary triangles the Dirichlet scan of
  Data_prepare.py:127-135 needs.
* `dirichlet_nodes`, `meshsize`, `stable_dt` restate Data_prepare.py:127-136 and
  Tools/commons.py:79-90 / Data_prepare.py:147 in vectorised form (same arithmetic per element,
  so the golden dt of Results/plotter.py:25 is reproduced bit-for-bit).
"""
from __future__ import annotations

import itertools

import numpy as np

_VTK_TETRA, _VTK_TRIANGLE = 10, 5


def read_vtk(path):
    """Legacy ASCII VTK UNSTRUCTURED_GRID -> (points f64 (N,3), tetra i64 (nE,4), triangles i64 (nF,3))."""
    with open(path) as f:
        tok = f.read().split()
    i = tok.index("POINTS")
    n = int(tok[i + 1])
    pts = np.array(tok[i + 3:i + 3 + 3 * n], dtype=np.float64).reshape(n, 3)
    i = tok.index("CELLS")
    nc, tot = int(tok[i + 1]), int(tok[i + 2])
    flat = np.array(tok[i + 3:i + 3 + tot], dtype=np.int64)
    j = tok.index("CELL_TYPES")
    types = np.array(tok[j + 2:j + 2 + nc], dtype=np.int64)
    # offsets of each cell record in `flat` (record = count followed by ids)
    counts = np.empty(nc, dtype=np.int64)
    p = 0
    starts = np.empty(nc, dtype=np.int64)
    for c in range(nc):
        starts[c] = p
        counts[c] = flat[p]
        p += 1 + counts[c]
    if p != tot:
        raise ValueError(f"{path}: CELLS block is inconsistent ({p} != {tot})")

    def pick(vtk_type, width):
        sel = starts[types == vtk_type]
        if sel.size == 0:
            return np.zeros((0, width), dtype=np.int64)
        return flat[sel[:, None] + 1 + np.arange(width)[None, :]]

    return pts, pick(_VTK_TETRA, 4), pick(_VTK_TRIANGLE, 3)


def write_vtk(path, points, cells, facets=None):
    """Write a legacy ASCII VTK file with the tetra (+ optional triangle) cells (round-trips read_vtk)."""
    facets = np.zeros((0, 3), dtype=np.int64) if facets is None else np.asarray(facets)
    with open(path, "w") as f:
        f.write("# vtk DataFile Version 2.0\nbeam, structured\nASCII\nDATASET UNSTRUCTURED_GRID\n")
        f.write(f"POINTS {len(points)} double\n")
        for p in points:
            f.write(f"{p[0]!r} {p[1]!r} {p[2]!r}\n".replace("np.float64(", "").replace(")", ""))
        nc = len(facets) + len(cells)
        f.write(f"\nCELLS {nc} {4 * len(facets) + 5 * len(cells)}\n")
        for t in facets:
            f.write(f"3 {t[0]} {t[1]} {t[2]}\n")
        for t in cells:
            f.write(f"4 {t[0]} {t[1]} {t[2]} {t[3]}\n")
        f.write(f"\nCELL_TYPES {nc}\n")
        f.write("5\n" * len(facets))
        f.write("10\n" * len(cells))


# Kuhn split of the unit cube: one tetrahedron per permutation of the axes, walking from corner
# (0,0,0) to (1,1,1).  Odd permutations are re-ordered so that det J > 0 for every element.
def _kuhn_corner_table():
    tets = []
    for perm in itertools.permutations(range(3)):
        v = [np.zeros(3, dtype=np.int64)]
        for ax in perm:
            w = v[-1].copy()
            w[ax] += 1
            v.append(w)
        v = np.array(v)
        J = (v[1:] - v[0]).T
        if np.linalg.det(J.astype(float)) < 0:
            v[[2, 3]] = v[[3, 2]]
        tets.append(v)
    return np.array(tets)  # (6, 4, 3) corner offsets


_KUHN = _kuhn_corner_table()


def structured_beam_dims(m, length=25):
    nx, ny, nz = length * m, m, m
    return nx, ny, nz


def structured_beam(m, length=25, dtype_index=np.int64, with_facets="x0", nx=None):
    """25 x 1 x 1 cantilever as (25m) x m x m hexahedra -> 6 Kuhn tets each.

    Node id is lexicographic with z fastest, then y, x slowest: id = (ix*(ny+1)+iy)*(nz+1)+iz, so
    x-slabs are contiguous id ranges.  Elements are numbered hex-major (hex id lexicographic the
    same way), 6 consecutive tets per hex.

    with_facets: "x0" -> only the boundary triangles on the clamped face x=0 (all that
    Data_prepare.py:127-135 looks at); "all" -> the whole boundary surface; None -> no facets.
    """
    nx_full, ny, nz = structured_beam_dims(m, length)
    nx = nx_full if nx is None else int(nx)          # nx: only the first nx hexahedron layers (a slab sample of the beam)
    h = 1.0 / m
    gx = np.arange(nx + 1, dtype=np.float64) * h
    gy = np.arange(ny + 1, dtype=np.float64) * h
    gz = np.arange(nz + 1, dtype=np.float64) * h
    X, Y, Z = np.meshgrid(gx, gy, gz, indexing="ij")
    points = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)

    def nid(ix, iy, iz):
        return (ix * (ny + 1) + iy) * (nz + 1) + iz

    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
    nhex = ix.size
    cells = np.empty((nhex, 6, 4), dtype=dtype_index)
    for t in range(6):
        for a in range(4):
            o = _KUHN[t, a]
            cells[:, t, a] = nid(ix + o[0], iy + o[1], iz + o[2])
    cells = cells.reshape(nhex * 6, 4)

    facets = None
    if with_facets is not None:
        fl = []

        def quad_face(fixed_axis, fixed_val):
            # the Kuhn diagonal of every hexahedron face joins the face's min and max corners
            axes = [a for a in range(3) if a != fixed_axis]
            n = [nx, ny, nz]
            ia, ib = np.meshgrid(np.arange(n[axes[0]]), np.arange(n[axes[1]]), indexing="ij")
            ia, ib = ia.ravel(), ib.ravel()

            def node(da, db):
                idx = [None, None, None]
                idx[fixed_axis] = np.full_like(ia, fixed_val)
                idx[axes[0]] = ia + da
                idx[axes[1]] = ib + db
                return nid(idx[0], idx[1], idx[2])

            a, b, c, d = node(0, 0), node(1, 0), node(1, 1), node(0, 1)
            return np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)])

        fl.append(quad_face(0, 0))
        if with_facets == "all":
            fl.append(quad_face(0, nx))
            fl.append(quad_face(1, 0))
            fl.append(quad_face(1, ny))
            fl.append(quad_face(2, 0))
            fl.append(quad_face(2, nz))
        facets = np.concatenate(fl).astype(dtype_index)
    return points, cells, facets


def dirichlet_nodes(points, facets, tol=1e-9):
    """Clamped nodes: every node of a boundary triangle whose 3 vertices have |x| < tol, in
    discovery order (Data_prepare.py:127-135)."""
    facets = np.asarray(facets)
    if facets.size == 0:
        return np.zeros(0, dtype=np.int64)
    on = np.all(np.abs(points[facets, 0]) < tol, axis=1)
    flat = facets[on].ravel()
    uniq, first = np.unique(flat, return_index=True)
    return uniq[np.argsort(first, kind="stable")].astype(np.int64)


def meshsize(cells, points):
    """CFL length 2*min_edge/sqrt(24) over the given tets (Tools/commons.py:79-90).

    np.linalg.norm of a 3-vector is sqrt(dot(x,x)); the same call is made here on stacked edge
    vectors so the per-edge values are the ones the reference computes.
    """
    P = points[np.asarray(cells)[:, :4]]
    pairs = ((0, 1), (1, 2), (2, 3), (1, 3), (0, 3), (0, 2))
    approx = []
    for a, b in pairs:
        e = P[:, a, :] - P[:, b, :]
        approx.append(np.sqrt(np.einsum("ij,ij->i", e, e)))
    approx = np.stack(approx, axis=1)                      # (nE, 6) screening values
    lo = approx.min()
    # the final value is re-evaluated with the very call the reference makes (np.linalg.norm on
    # the 3-vector) for every edge within rounding distance of the minimum, so the result is
    # bit-identical to commons.py:79-90 whatever summation order einsum used above
    vecs = []
    for k, (a, b) in enumerate(pairs):
        sel = approx[:, k] <= lo * (1 + 1e-12)
        if sel.any():
            vecs.append(P[sel, a, :] - P[sel, b, :])
    vecs = np.unique(np.concatenate(vecs), axis=0)         # distinct candidate edge vectors
    best = min(np.linalg.norm(v) for v in vecs)
    return 2.0 * best / np.sqrt(24)


def stable_dt(cells, points, E=1e6, nu=0.3, rho=1, gamma=.9):
    """dt = gamma * Meshsize / sqrt(E/rho/(1-nu**2))  (Data_prepare.py:147)."""
    return gamma * meshsize(cells, points) / np.sqrt(E / rho / (1 - nu ** 2))
