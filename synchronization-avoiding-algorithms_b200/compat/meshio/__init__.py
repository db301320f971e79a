"""Stand-in for meshio (legacy ASCII VTK only): see compat/README.md."""
import _saa_defer

_real = _saa_defer.real("meshio", __file__)
if _real is not None:
    import sys as _sys
    _sys.modules[__name__] = _real
else:
    import numpy as np

    import _saa_bootstrap  # noqa: F401

    from saa_b200 import mesh as _mesh


    class _Block:
        def __init__(self, type_, data):
            self.type, self.data = type_, data


    class Mesh:
        def __init__(self, points, tetra, triangle):
            self.points = points
            self.cells_dict = {"tetra": tetra, "triangle": triangle}
            self.cells = [_Block("triangle", triangle), _Block("tetra", tetra)]


    def read(path):
        p, c, f = _mesh.read_vtk(str(path))
        return Mesh(p, c, f)


    def write_points_cells(path, points, cells, point_data=None):
        tet = [b.data for b in cells if getattr(b, "type", None) == "tetra"]
        tri = [b.data for b in cells if getattr(b, "type", None) == "triangle"]
        _mesh.write_vtk(str(path), np.asarray(points), tet[0] if tet else np.zeros((0, 4), dtype=np.int64),
                        tri[0] if tri else None)
        if point_data:
            with open(str(path), "a") as fh:
                fh.write(f"\nPOINT_DATA {len(points)}\n")
                for name, v in point_data.items():
                    fh.write(f"SCALARS {name} double 1\nLOOKUP_TABLE default\n")
                    fh.write("\n".join(repr(float(x)) for x in np.asarray(v).reshape(-1)) + "\n")
