"""Set-up at scale on the GPU: per-rank mesh slab, local numbering, sparse assembly (K6), halo description and
device plan — nothing of O(mesh) ever lives in host memory except the (sorted) node lists used to find the
partition interfaces.

The reference's set-up (/root/reference/Data_prepare.py:104-209, Tools/Mat_construction.py:122-231) is O(N^2)
in memory and cannot run beyond ~3e4 DOF; this module produces the same quantities — Local_nodal_list in
first-appearance order, LocalK as CSR with ascending columns and exact zeros dropped, lumped mass, un-ramped
load, clamped DOFs, dt — for 1e6..1e8 DOF.  Element matrices are evaluated in closed form on the device, so
entries agree with the reference's to a few 1e-16 relative (tests/test_gpu_device_setup.py), not bit for bit;
the time-step kernels downstream are the same bit-exact ones.

torch is used for device memory and sort/unique plumbing; the numeric kernels are the library's.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import maps, mesh
from .plan import SaaError, StepPlan, _check, lib
from .problem import DAMP_DEFAULT, E_DEFAULT, FZ_DEFAULT, GAMMA_DEFAULT, NU_DEFAULT, RHO_DEFAULT, lame


# ---- structured cantilever, one x-slab per rank ------------------------------------------------------------
_LAYER_WEIGHTS = None     # optional relative speeds of the ranks (set_layer_weights): faster GPUs get more layers


def set_layer_weights(weights):
    """Relative throughput of every rank (e.g. measured DOF-steps/s of a short local run).  None = equal slabs."""
    global _LAYER_WEIGHTS
    _LAYER_WEIGHTS = None if weights is None else [float(w) for w in weights]


def layer_bounds(m, size, length=25):
    """Hexahedron layers (index along x) of each rank: rank r owns layers [b[r], b[r+1])."""
    nx = length * m
    if _LAYER_WEIGHTS is not None and len(_LAYER_WEIGHTS) == size:
        w = np.asarray(_LAYER_WEIGHTS) / np.sum(_LAYER_WEIGHTS)
        b = np.rint(np.concatenate([[0.0], np.cumsum(w)]) * nx).astype(int)
        b[0], b[-1] = 0, nx
        for r in range(1, size + 1):              # every rank keeps at least one layer
            b[r] = max(b[r], b[r - 1] + 1)
        b[-1] = nx
        return [int(x) for x in b]
    return [(r * nx) // size for r in range(size + 1)]


def layer_slab_partition(m, size, length=25):
    """epart (host, int64) of mesh.structured_beam(m) for the layer-slab partition — for tests / small meshes."""
    nx, ny, nz = mesh.structured_beam_dims(m, length)
    b = np.asarray(layer_bounds(m, size, length))
    layer = np.arange(nx * ny * nz * 6, dtype=np.int64) // (6 * ny * nz)
    return np.searchsorted(b, layer, side="right") - 1


def structured_slab_cells(m, rank, size, length=25, device="cuda"):
    """Global-id connectivity (nE_loc,4) int64 of this rank's elements of mesh.structured_beam(m), in ascending
    global element order, generated on the device."""
    import torch
    nx, ny, nz = mesh.structured_beam_dims(m, length)
    b = layer_bounds(m, size, length)
    x0, x1 = b[rank], b[rank + 1]
    kuhn = torch.as_tensor(mesh._KUHN, device=device)                     # (6,4,3) corner offsets
    ix = torch.arange(x0, x1, device=device, dtype=torch.int64)
    iy = torch.arange(ny, device=device, dtype=torch.int64)
    iz = torch.arange(nz, device=device, dtype=torch.int64)
    IX, IY, IZ = torch.meshgrid(ix, iy, iz, indexing="ij")
    IX, IY, IZ = IX.reshape(-1, 1, 1), IY.reshape(-1, 1, 1), IZ.reshape(-1, 1, 1)
    nid = ((IX + kuhn[None, :, :, 0]) * (ny + 1) + (IY + kuhn[None, :, :, 1])) * (nz + 1) + (IZ + kuhn[None, :, :, 2])
    return nid.reshape(-1, 4)


def block_grid(size):
    """Process grid (px, py, pz) of the block partition: cuts along x first, then y, then z (8 -> 2 x 2 x 2, so that a
    rank has up to 7 neighbours and edges / the centre line of the beam are held by 4 / 8 ranks)."""
    g = [1, 1, 1]
    ax, left = 0, int(size)
    while left > 1:
        if left % 2:
            raise ValueError("block partition needs a power-of-two number of ranks")
        g[ax % 3] *= 2
        ax += 1
        left //= 2
    return tuple(g)


def structured_block_cells(m, rank, size, length=25, grid=None, device="cuda"):
    """Same as structured_slab_cells for a px x py x pz block partition of the hexahedra (rank = (bx*py + by)*pz + bz):
    this rank's elements in ascending global element order."""
    import torch
    nx, ny, nz = mesh.structured_beam_dims(m, length)
    px, py, pz = grid or block_grid(size)
    if px * py * pz != size:
        raise ValueError(f"block grid {px}x{py}x{pz} does not match {size} ranks")
    bz, by, bx = rank % pz, (rank // pz) % py, rank // (pz * py)
    rng = lambda n, p, b: ((b * n) // p, ((b + 1) * n) // p)
    (x0, x1), (y0, y1), (z0, z1) = rng(nx, px, bx), rng(ny, py, by), rng(nz, pz, bz)
    kuhn = torch.as_tensor(mesh._KUHN, device=device)
    ix = torch.arange(x0, x1, device=device, dtype=torch.int64)
    iy = torch.arange(y0, y1, device=device, dtype=torch.int64)
    iz = torch.arange(z0, z1, device=device, dtype=torch.int64)
    IX, IY, IZ = torch.meshgrid(ix, iy, iz, indexing="ij")               # lexicographic = ascending global hexahedron id
    IX, IY, IZ = IX.reshape(-1, 1, 1), IY.reshape(-1, 1, 1), IZ.reshape(-1, 1, 1)
    nid = ((IX + kuhn[None, :, :, 0]) * (ny + 1) + (IY + kuhn[None, :, :, 1])) * (nz + 1) + (IZ + kuhn[None, :, :, 2])
    return nid.reshape(-1, 4)


def block_partition(m, size, length=25, grid=None):
    """epart (host, int64) of mesh.structured_beam(m) for the block partition — for tests / small meshes."""
    nx, ny, nz = mesh.structured_beam_dims(m, length)
    px, py, pz = grid or block_grid(size)
    h = np.arange(nx * ny * nz, dtype=np.int64)
    iz, iy, ix = h % nz, (h // nz) % ny, h // (nz * ny)
    own = lambda i, n, p: np.minimum(((i + 1) * p - 1) // n, p - 1)    # block b owns [b*n//p, (b+1)*n//p)
    r = (own(ix, nx, px) * py + own(iy, ny, py)) * pz + own(iz, nz, pz)
    return np.repeat(r, 6)


def structured_points(m, node_ids, length=25):
    """Coordinates (n,3) float64 of the given global node ids of mesh.structured_beam(m) (device tensor)."""
    import torch
    nx, ny, nz = mesh.structured_beam_dims(m, length)
    h = 1.0 / m
    iz = node_ids % (nz + 1)
    iy = (node_ids // (nz + 1)) % (ny + 1)
    ix = node_ids // ((nz + 1) * (ny + 1))
    return torch.stack([ix.double() * h, iy.double() * h, iz.double() * h], dim=1).contiguous()


# ---- local numbering ------------------------------------------------------------------------------------------
def local_numbering(cells_global):
    """(Local_nodal_list, cells in local ids) for one rank's elements — Distributed_tools.py:14-24 semantics:
    nodes in first-appearance order while scanning the elements row by row.  Device tensors in and out."""
    import torch
    flat = cells_global.reshape(-1)
    uniq, inv = torch.unique(flat, return_inverse=True)                 # sorted ids, position of every entry
    first = torch.full((uniq.numel(),), flat.numel(), dtype=torch.int64, device=flat.device)
    first.scatter_reduce_(0, inv, torch.arange(flat.numel(), device=flat.device), reduce="amin")
    order = torch.argsort(first)                                        # first-appearance order
    rank_of = torch.empty_like(order)
    rank_of[order] = torch.arange(order.numel(), device=flat.device)
    cells_loc = rank_of[inv].reshape(-1, 4).to(torch.int32).contiguous()
    return uniq[order].contiguous(), cells_loc


class DeviceCSR:
    """CSR arrays allocated by saa_assemble_stiffness_dev (int64 indptr, int32 indices, float64 data)."""

    def __init__(self, n_rows, nnz, indptr, indices, data):
        self.n_rows, self.nnz, self.indptr, self.indices, self.data = n_rows, nnz, indptr, indices, data

    def to_scipy(self):
        from scipy.sparse import csr_matrix
        ip = np.empty(self.n_rows + 1, dtype=np.int64)
        ix = np.empty(self.nnz, dtype=np.int32)
        dv = np.empty(self.nnz, dtype=np.float64)
        L = lib()
        for dst, src in ((ip, self.indptr), (ix, self.indices), (dv, self.data)):
            _check(L.saa_device_copy(dst.ctypes.data_as(ctypes.c_void_p), src, dst.nbytes), "saa_device_copy")
        K = csr_matrix((dv, ix, ip.astype(np.int32) if self.nnz < 2 ** 31 else ip), shape=(self.n_rows, self.n_rows))
        K.has_sorted_indices = True
        return K

    def free(self):
        L = lib()
        for a in ("indptr", "indices", "data"):
            p = getattr(self, a)
            if p:
                L.saa_device_free(p)
                setattr(self, a, None)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def assemble_stiffness(cells_loc, points_loc, lmd, mu, device=0):
    """K6: LocalK of one rank on the device -> DeviceCSR (Mat_construction.py:122-150 semantics)."""
    n_nodes, n_elem = points_loc.shape[0], cells_loc.shape[0]
    ip, ix, dv = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    nnz = ctypes.c_int64(0)
    _check(lib().saa_assemble_stiffness_dev(device, n_nodes, n_elem, cells_loc.data_ptr(), points_loc.data_ptr(), float(lmd),
                                            float(mu), ctypes.byref(ip), ctypes.byref(ix), ctypes.byref(dv), ctypes.byref(nnz)),
           "saa_assemble_stiffness_dev")
    return DeviceCSR(3 * n_nodes, nnz.value, ip, ix, dv)


def assemble_mass_load(cells_loc, points_loc, rho, fz, device=0):
    """Partial (local elements only) lumped mass per node (n,) and load vector (3n,) as device tensors."""
    import torch
    n_nodes, n_elem = points_loc.shape[0], cells_loc.shape[0]
    m = torch.empty(n_nodes, dtype=torch.float64, device=points_loc.device)
    F = torch.empty(3 * n_nodes, dtype=torch.float64, device=points_loc.device)
    _check(lib().saa_assemble_mass_load_dev(device, n_nodes, n_elem, cells_loc.data_ptr(), points_loc.data_ptr(), float(rho),
                                            float(fz), m.data_ptr(), F.data_ptr()), "saa_assemble_mass_load_dev")
    return m, F


def min_edge_meshsize(cells_loc, points_loc, chunk=1 << 23):
    """Tools/commons.py:79-90 at scale: screening of the minimum edge on the device (in chunks of elements), the
    final value with the reference's own call (np.linalg.norm on the candidate edge vectors) on the host."""
    import torch
    pairs = ((0, 1), (1, 2), (2, 3), (1, 3), (0, 3), (0, 2))
    nE = cells_loc.shape[0]

    def edges(c0):
        P = points_loc[cells_loc[c0:c0 + chunk].long()]                 # (chunk,4,3)
        for a, b in pairs:
            e = P[:, a, :] - P[:, b, :]
            yield e, torch.sqrt((e * e).sum(1))

    lo = float("inf")
    for c0 in range(0, nE, chunk):
        for _, l in edges(c0):
            lo = min(lo, float(l.min()))
    cand = []
    for c0 in range(0, nE, chunk):
        for e, l in edges(c0):
            sel = l <= lo * (1 + 1e-12)
            if bool(sel.any()):
                cand.append(torch.unique(e[sel], dim=0).cpu().numpy())
    cand = np.unique(np.concatenate(cand), axis=0)
    best = min(np.linalg.norm(v) for v in cand)
    return 2.0 * best / np.sqrt(24)


# ---- shared-node sums at set-up time ------------------------------------------------------------------------------
def holders_send(halo, own):
    """Messages {neighbour: rows of `own` (s,k) it shares with this rank, ascending global id}."""
    import torch
    return {nb: own[torch.as_tensor(halo["send_idx"][nb], device=own.device)].contiguous() for nb in halo["neighbours"]}


def holders_sum(halo, me, own, recv):
    """Values of the shared nodes summed over their holders in ascending rank order (the association of
    syn_cpus, Distributed_tools.py:83-86), starting from 0.0.  own: (s,k) partial values of this rank at
    halo['shared_pos']; recv: {neighbour: its holders_send message for this rank}."""
    import torch
    acc = torch.zeros_like(own)
    for r in sorted(list(halo["neighbours"]) + [me]):
        if r == me:
            acc = acc + own
        else:
            idx = torch.as_tensor(halo["send_idx"][r], device=own.device)
            acc[idx] = acc[idx] + recv[r]
    return acc


def dist_exchange(send):
    """Neighbour exchange of device tensors over the default torch.distributed group (NCCL; with a gloo group — ranks
    sharing a GPU in the tests — the set-up messages pass through host memory)."""
    import torch
    import torch.distributed as dist
    via_host = dist.get_backend() == "gloo"
    out = {nb: (t.cpu().contiguous() if via_host else t) for nb, t in send.items()}
    recv = {nb: torch.empty_like(t) for nb, t in out.items()}
    ops = []
    for nb in sorted(out):
        ops.append(dist.P2POp(dist.isend, out[nb], nb))
        ops.append(dist.P2POp(dist.irecv, recv[nb], nb))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return {nb: t.to(send[nb].device) for nb, t in recv.items()} if via_host else recv


def gather_node_lists(local_nodes, size):
    """All ranks' Local_nodal_list (host int64 arrays) — the input of maps.halo_plan."""
    import torch
    import torch.distributed as dist
    if size == 1:
        return [local_nodes.cpu().numpy()]
    dev = torch.device("cpu") if dist.get_backend() == "gloo" else local_nodes.device
    n = torch.tensor([local_nodes.numel()], device=dev)
    sizes = [torch.zeros_like(n) for _ in range(size)]
    dist.all_gather(sizes, n)
    mx = int(max(int(s) for s in sizes))
    buf = torch.zeros(mx, dtype=torch.int64, device=dev)
    buf[:local_nodes.numel()] = local_nodes
    out = [torch.empty_like(buf) for _ in range(size)]
    dist.all_gather(out, buf)
    return [o[:int(s)].cpu().numpy() for o, s in zip(out, sizes)]


# ---- one rank of a structured cantilever, end to end --------------------------------------------------------------
def rank_local(cells_global, points_of, clamped_of, rank, size, device_index=0, E=E_DEFAULT, nu=NU_DEFAULT, rho=RHO_DEFAULT,
               fz=FZ_DEFAULT, gamma=GAMMA_DEFAULT, n_global_nodes=0, n_global_elem=0, reorder=None, keep_mesh=False):
    """Phase 1 (no communication) for ANY tetrahedral mesh: `cells_global` (nE_loc,4) int64 device tensor with the
    global node ids of this rank's elements in ascending element order (Local_ele_list order); points_of(ids) ->
    (n,3) float64 device coordinates of the given global node ids; clamped_of(ids) -> bool device mask of clamped
    nodes.  Produces local numbering, stiffness, partial mass / load, local dt, clamped DOFs."""
    import torch
    lmd, mu = lame(E, nu)
    local_nodes, cells_loc = local_numbering(cells_global)
    pts = points_of(local_nodes)
    K = assemble_stiffness(cells_loc, pts, lmd, mu, device_index)
    m_node, F = assemble_mass_load(cells_loc, pts, rho, fz, device_index)
    dt_loc = gamma * min_edge_meshsize(cells_loc, pts) / np.sqrt(E / rho / (1 - nu ** 2))     # Data_prepare.py:147
    clamped = torch.nonzero(clamped_of(local_nodes)).reshape(-1).cpu().numpy()                 # (:127-144) ascending local position
    node_order = morton_node_order(pts) if reorder == "morton" else None
    return dict(cells_loc=cells_loc, pts=pts, lame=(lmd, mu), keep_mesh=keep_mesh,
                node_order=node_order, rank=rank, size=size, device_index=device_index, local_nodes=local_nodes, n_nodes=local_nodes.numel(),
                n_elem=cells_loc.shape[0], K=K, m_node=m_node, F=F, dt_loc=float(dt_loc),
                dirichlet=maps.node_to_dof(3, [0, 1, 2], clamped), n_global_nodes=n_global_nodes, n_global_elem=n_global_elem)


def structured_rank_local(m, rank, size, device_index=0, length=25, partition="slabs", grid=None, **kw):
    """Phase 1 for the structured cantilever, generated on the device: x-slab `rank` of `size` (partition="slabs")
    or block `rank` of a px x py x pz grid (partition="blocks")."""
    import torch
    dev = torch.device("cuda", device_index)
    nx, ny, nz = mesh.structured_beam_dims(m, length)
    if partition == "blocks":
        cells_g = structured_block_cells(m, rank, size, length, grid, device=dev)
    elif partition == "slabs":
        cells_g = structured_slab_cells(m, rank, size, length, device=dev)
    else:
        raise ValueError(f"unknown structured partition {partition!r}")
    return rank_local(cells_g, lambda ids: structured_points(m, ids, length), lambda ids: ids < (ny + 1) * (nz + 1), rank, size,
                      device_index, n_global_nodes=(nx + 1) * (ny + 1) * (nz + 1), n_global_elem=6 * nx * ny * nz, **kw)


def mesh_rank_local(points, cells, facets, epart, rank, size, device_index=0, **kw):
    """Phase 1 for a mesh given as host arrays (meshio / gmsh output) and an element -> rank vector `epart`:
    only this rank's elements and the coordinate table travel to the device.  The rows are laid out in HBM along a
    Morton curve of the node coordinates by default (reorder="morton"): the element order of a mesh file need not be
    spatially coherent, and a random one costs 36 % at 21 M DOF (profiles/r1/locality.md); reorder=None keeps the
    first-appearance order."""
    kw.setdefault("reorder", "morton")
    import torch
    dev = torch.device("cuda", device_index)
    cells = np.asarray(cells, dtype=np.int64)
    ele = np.nonzero(np.asarray(epart) == rank)[0]
    cells_g = torch.from_numpy(cells[ele]).to(dev)
    P = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64)).to(dev)
    is_clamped = torch.zeros(len(points), dtype=torch.bool, device=dev)
    D = mesh.dirichlet_nodes(np.asarray(points), np.asarray(facets))                           # Data_prepare.py:127-135
    if len(D):
        is_clamped[torch.from_numpy(D).to(dev)] = True
    return rank_local(cells_g, lambda ids: P[ids].contiguous(), lambda ids: is_clamped[ids], rank, size, device_index,
                      n_global_nodes=len(points), n_global_elem=len(cells), **kw)


rank_plan = None   # set below (generic name of structured_rank_plan)


def shared_partials(loc, halo):
    """(s,4) partial [mass, Fx, Fy, Fz] of this rank at its shared nodes."""
    import torch
    sp = torch.as_tensor(halo["shared_pos"], device=loc["m_node"].device)
    return torch.cat([loc["m_node"][sp, None], loc["F"].view(-1, 3)[sp]], dim=1)


def morton_node_order(pts):
    """Z-order (Morton) permutation of the local nodes from their coordinates: nodes close in space end up close in
    HBM whatever the element order of the mesh file was.  21 bits per axis."""
    import torch
    lo, hi = pts.min(0).values, pts.max(0).values
    q = ((pts - lo) / torch.clamp(hi - lo, min=1e-300) * (2 ** 21 - 1)).long().clamp_(0, 2 ** 21 - 1)

    def spread(v):                                          # insert two zero bits between the bits of v
        v = (v | (v << 32)) & 0x1F00000000FFFF
        v = (v | (v << 16)) & 0x1F0000FF0000FF
        v = (v | (v << 8)) & 0x100F00F00F00F00F
        v = (v | (v << 4)) & 0x10C30C30C30C30C3
        v = (v | (v << 2)) & 0x1249249249249249
        return v
    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    return torch.argsort(code, stable=True).to(torch.int32).cpu().numpy()


def structured_rank_plan(loc, halo, recv, dt, alpha=DAMP_DEFAULT, keep_csr=False):
    """Phase 2: fold the neighbours' partial mass / load in (rank-ordered sums), build the device plan."""
    import torch
    rank, size = loc["rank"], loc["size"]
    if size > 1 and len(halo["shared_pos"]):
        sp = torch.as_tensor(halo["shared_pos"], device=loc["m_node"].device)
        tot = holders_sum(halo, rank, shared_partials(loc, halo), recv)
        loc["m_node"][sp] = tot[:, 0]
        loc["F"].view(-1, 3)[sp] = tot[:, 1:]
    lM = loc["m_node"].repeat_interleave(3).contiguous()
    torch.cuda.synchronize()
    K = loc["K"]
    pl = StepPlan.from_device(3 * loc["n_nodes"], K.indptr, K.indices, K.data, loc["F"].data_ptr(), lM.data_ptr(),
                              loc["dirichlet"], dt, alpha, device=loc["device_index"], halo=halo if size > 1 else None,
                              rank=rank, size=size, node_order=loc.get("node_order"))
    info = dict(n_nodes=loc["n_nodes"], n_elem=loc["n_elem"], nnz=K.nnz, dt=dt, local_nodes=loc["local_nodes"], halo=halo,
                n_global_nodes=loc["n_global_nodes"], n_global_elem=loc["n_global_elem"])
    if loc.get("keep_mesh"):                       # local connectivity / coordinates for the matrix-free kernel (StepPlan.set_matfree)
        info.update(cells_loc=loc["cells_loc"], pts=loc["pts"], lame=loc["lame"])
    else:
        loc.pop("cells_loc", None); loc.pop("pts", None)
    if keep_csr:
        info.update(K=K, F=loc["F"], lM=lM, dirichlet=loc["dirichlet"])
    else:
        K.free()
    return pl, info


def build_structured_rank(m, rank, size, device_index=0, alpha=DAMP_DEFAULT, keep_csr=False, **kw):
    """One process per GPU (collective over torch.distributed when size > 1): (StepPlan, info)."""
    return finish_rank(structured_rank_local(m, rank, size, device_index, **kw), alpha, keep_csr)


def build_mesh_rank(points, cells, facets, epart, rank, size, device_index=0, alpha=DAMP_DEFAULT, keep_csr=False, **kw):
    """Same for an arbitrary tetrahedral mesh + partition vector (one process per GPU, collective when size > 1)."""
    return finish_rank(mesh_rank_local(points, cells, facets, epart, rank, size, device_index, **kw), alpha, keep_csr)


def finish_rank(loc, alpha=DAMP_DEFAULT, keep_csr=False):
    """Phase 2 of a one-process-per-GPU build: interface description, rank-ordered mass / load sums, global dt, plan."""
    import torch
    rank, size = loc["rank"], loc["size"]
    halo, recv, dt = None, None, loc["dt_loc"]
    if size > 1:
        import torch.distributed as dist
        lists = gather_node_lists(loc["local_nodes"], size)
        halo = maps.halo_plan(rank, size, lists)
        recv = dist_exchange(holders_send(halo, shared_partials(loc, halo)))
        t = torch.tensor([dt], dtype=torch.float64, device="cpu" if dist.get_backend() == "gloo" else loc["m_node"].device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)                                               # Data_prepare.py:151-154
        dt = float(t.item())
    out = structured_rank_plan(loc, halo, recv, dt, alpha, keep_csr)
    torch.cuda.empty_cache()
    return out


def build_structured_in_process(m, size, device_index=0, alpha=DAMP_DEFAULT, keep_csr=False, **kw):
    """All ranks of the layer-slab partition in ONE process on one GPU (tests; P partitions on fewer GPUs):
    returns (plans, infos); wrap the plans in plan.PlanGroup to step them together."""
    return finish_in_process([structured_rank_local(m, r, size, device_index, **kw) for r in range(size)], alpha, keep_csr)


def build_mesh_in_process(points, cells, facets, epart, size, device_index=0, alpha=DAMP_DEFAULT, keep_csr=False, **kw):
    """All ranks of an arbitrary partitioned mesh in ONE process on one GPU: (plans, infos)."""
    return finish_in_process([mesh_rank_local(points, cells, facets, epart, r, size, device_index, **kw) for r in range(size)],
                             alpha, keep_csr)


def finish_in_process(locs, alpha=DAMP_DEFAULT, keep_csr=False):
    size = len(locs)
    lists = [l["local_nodes"].cpu().numpy() for l in locs]
    halos = [maps.halo_plan(r, size, lists) if size > 1 else None for r in range(size)]
    sends = [holders_send(halos[r], shared_partials(locs[r], halos[r])) if size > 1 else {} for r in range(size)]
    dt = min(l["dt_loc"] for l in locs)
    out = []
    for r in range(size):
        recv = {nb: sends[nb][r] for nb in halos[r]["neighbours"]} if size > 1 else None
        out.append(structured_rank_plan(locs[r], halos[r], recv, dt, alpha, keep_csr))
    return [o[0] for o in out], [o[1] for o in out]


rank_plan = structured_rank_plan
