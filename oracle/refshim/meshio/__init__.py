"""Minimal legacy-ASCII-VTK reader standing in for meshio.read (oracle harness only).

Returns an object with `.points` (N,3) float64 and `.cells_dict` with the
'tetra' (VTK type 10) and 'triangle' (VTK type 5) connectivity as int64 arrays,
which is all /root/reference/Data_prepare.py:58-61 uses.
"""
import numpy as np

_VTK_NAMES = {1: "vertex", 3: "line", 5: "triangle", 10: "tetra"}


class _Mesh:
    def __init__(self, points, cells_dict):
        self.points = points
        self.cells_dict = cells_dict
        self.cells = [(k, v) for k, v in cells_dict.items()]


def read(path):
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
    cells = {}
    p = 0
    for t in types:
        k = int(flat[p])
        cells.setdefault(_VTK_NAMES.get(int(t), str(int(t))), []).append(flat[p + 1:p + 1 + k])
        p += 1 + k
    return _Mesh(pts, {k: np.array(v, dtype=np.int64) for k, v in cells.items()})


def write_points_cells(*a, **k):  # Data_prepare.py:168 (steady output; unused by the harness)
    pass
