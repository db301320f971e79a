"""`part_mesh_kway(nparts, eptr, eind)` (the call of Data_prepare.py:94): every rank passes its own contiguous
chunk of elements; the chunks are gathered, partitioned by serial METIS (dual graph, tets sharing a face), and
each rank gets the part vector of its own chunk back, like ParMETIS_V3_PartMeshKway."""
import numpy as np

import _saa_bootstrap  # noqa: F401

from saa_b200 import comm as _comm
from saa_b200 import partition as _partition


def part_mesh_kway(nparts, eptr, eind, **_):
    c = _comm.world()
    eptr = np.asarray(eptr, dtype=np.int64)
    eind = np.asarray(eind, dtype=np.int64)
    if np.any(np.diff(eptr) != 4):
        raise NotImplementedError("tetrahedral meshes only")
    chunks = c.allgather(eind.reshape(-1, 4))
    cells = np.concatenate(chunks)
    epart = _partition.metis_part_mesh(cells, int(cells.max()) + 1, int(nparts))
    start = sum(len(x) for x in chunks[:c.Get_rank()])
    return 0, epart[start:start + len(chunks[c.Get_rank()])]
