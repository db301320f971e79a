"""CPU checks of the host-side plumbing around the device path: slab generator and numbering of the device set-up
(torch ops, run here on CPU tensors), layer balancing, Morton order, the serial communicator, the driver launcher."""
import os
import subprocess
import sys

import numpy as np
import pytest

import saa_b200  # noqa: F401
from saa_b200 import comm, device_setup as ds, maps, mesh, plan as splan
from util import ROOT

PKG = os.path.join(ROOT, "synchronization-avoiding-algorithms_b200")


@pytest.mark.parametrize("m,size", [(2, 1), (2, 3), (3, 4)])
def test_slab_generator_and_numbering_equal_host_maps(m, size):
    import torch
    pts, cells, fac = mesh.structured_beam(m)
    ep = ds.layer_slab_partition(m, size)
    assert ep.min() == 0 and ep.max() == size - 1 and np.all(np.diff(ep) >= 0)
    for r in range(size):
        c = ds.structured_slab_cells(m, r, size, device="cpu")
        assert np.array_equal(c.numpy(), cells[ep == r])                         # ascending global element order
        nodes, cl = ds.local_numbering(c)
        ele, ref_nodes = maps.rankwise_dist(r, ep, cells)
        assert np.array_equal(nodes.numpy(), ref_nodes)                           # first-appearance order, exact
        assert cl.dtype == torch.int32 and np.array_equal(nodes.numpy()[cl.numpy()], cells[ep == r])
        P = ds.structured_points(m, nodes)
        assert np.array_equal(P.numpy(), pts[ref_nodes])
        assert ds.min_edge_meshsize(cl, P, chunk=50) == mesh.meshsize(cells[ep == r], pts)


def test_layer_bounds_follow_weights():
    try:
        assert ds.layer_bounds(65, 8) == [(r * 1625) // 8 for r in range(9)]
        ds.set_layer_weights([1, 1, 0.5, 1])
        b = ds.layer_bounds(4, 4)
        w = np.diff(b)
        assert b[0] == 0 and b[-1] == 100 and w[2] < w[0] and abs(w[2] - 100 * 0.5 / 3.5) <= 1 and w.min() >= 1
        ds.set_layer_weights([1e-9, 1, 1])
        assert np.diff(ds.layer_bounds(1, 3)).min() >= 1                          # nobody ends up without a layer
    finally:
        ds.set_layer_weights(None)


def test_morton_order_is_a_local_permutation():
    import torch
    rng = np.random.default_rng(0)
    pts = torch.from_numpy(rng.random((4096, 3)))
    o = ds.morton_node_order(pts)
    assert sorted(o.tolist()) == list(range(4096))
    p = pts.numpy()
    jump_sorted = np.linalg.norm(np.diff(p[o], axis=0), axis=1).mean()
    jump_given = np.linalg.norm(np.diff(p, axis=0), axis=1).mean()
    assert jump_sorted < 0.3 * jump_given
    # a structured grid in lexicographic order: the Z-curve visits every node once
    g = torch.from_numpy(np.stack(np.meshgrid(np.arange(4.), np.arange(4.), np.arange(4.), indexing="ij"), -1).reshape(-1, 3))
    assert sorted(ds.morton_node_order(g).tolist()) == list(range(64))


def test_serial_communicator_has_the_mpi4py_surface_the_drivers_use():
    c = comm.SerialComm()
    assert (c.Get_rank(), c.Get_size()) == (0, 1)
    assert c.bcast({"a": 1}, root=0) == {"a": 1} and c.gather(5, root=0) == [5] and c.allgather("x") == ["x"]
    buf = np.empty(3)
    c.Gatherv(np.arange(3.0), buf, root=0)
    assert np.array_equal(buf, np.arange(3.0))
    c.Gather(np.float64(2.5), buf[:1], root=0)
    assert buf[0] == 2.5
    c.Barrier()
    assert c.exchange(np.zeros(0), np.zeros(0, dtype=np.int32), np.zeros(1, dtype=np.int64)).size == 0


def test_run_driver_puts_the_package_first_and_compat_last(tmp_path):
    """A script living next to a decoy `Tools` directory must still get the drop-in `Tools`."""
    os.makedirs(tmp_path / "Tools")
    (tmp_path / "Tools" / "__init__.py").write_text("DECOY = True\n")
    (tmp_path / "Tools" / "commons.py").write_text("raise RuntimeError('decoy Tools imported')\n")
    (tmp_path / "drv.py").write_text(
        "from Tools.commons import *\nimport Tools, sys, h5py, meshio\n"
        "assert not hasattr(Tools, 'DECOY')\nassert linear_ramp(0.25) == 0.25 and linear_ramp(3) == 1.0\n"
        "assert sys.argv[1:] == ['--flag', '7']\nprint('driver ok', node_to_dof(3, [0, 1, 2], [2]).tolist())\n")
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    r = subprocess.run([sys.executable, os.path.join(PKG, "run_driver.py"), str(tmp_path / "drv.py"), "--flag", "7"], cwd=str(tmp_path),
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "driver ok [6, 7, 8]" in r.stdout


def test_rcm_order_is_a_permutation_that_reduces_bandwidth():
    from saa_b200 import assembly, plan, problem
    pts, cells, _ = mesh.structured_beam(2)
    rng = np.random.default_rng(1)
    perm = rng.permutation(len(pts))                                             # scramble the numbering
    K = assembly.local_stiffness_csr(perm, cells, pts, *problem.lame(1e6, 0.3))
    o = plan.rcm_node_order(K)
    assert sorted(o.tolist()) == list(range(len(pts)))
    G = (K != 0).tocoo()
    pos = np.empty(len(pts), dtype=np.int64)
    pos[o] = np.arange(len(pts))
    bw_rcm = np.abs(pos[G.row // 3] - pos[G.col // 3]).max()
    bw_raw = np.abs(G.row // 3 - G.col // 3).max()
    assert bw_rcm < 0.5 * bw_raw


def test_bench_reference_arm_prints_one_json_line_without_a_gpu():
    """`bench.py --impl reference` (the CPU restatement timed on the host cores) needs no GPU and prints exactly one
    JSON line with the contract's keys."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--refine", "3", "--cpu-seconds", "0.3"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "DOF-steps/sec" and d["unit"] == "DOF-steps/s" and d["higher_is_better"]
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]
    # ranks other than 0 of a torchrun launch do nothing and exit 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_compat_stand_ins_step_aside_for_real_packages(tmp_path):
    """compat/ may sit anywhere on PYTHONPATH (even first): a real h5py / mpi4py / meshio found elsewhere wins, so the
    result files come from h5py whenever it is installed; without it the stand-in writes genuine HDF5 itself (hdf5_lite)."""
    import subprocess
    import sys
    from util import ROOT
    pkg = os.path.join(ROOT, "synchronization-avoiding-algorithms_b200")
    fake = tmp_path / "site"
    for name in ("h5py", "meshio"):
        (fake / name).mkdir(parents=True)
        (fake / name / "__init__.py").write_text("IS_REAL = True\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(pkg, "compat"), str(fake)]))
    code = ("import h5py, meshio, mpi4py\nfrom mpi4py import MPI\n"
            "assert h5py.IS_REAL and meshio.IS_REAL and not hasattr(h5py, 'IS_STAND_IN')\n"
            "assert MPI.COMM_WORLD.Get_size() == 1\nprint('ok')")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
    # no real h5py: the stand-in, writing HDF5 at the very path the caller names
    env = dict(os.environ, PYTHONPATH=os.path.join(pkg, "compat"))
    code = ("import h5py, numpy as np, os\nf = h5py.File('x.hdf5', 'w'); f.create_dataset('Displacement', data=np.eye(2), compression='gzip'); f.close()\n"
            "assert h5py.IS_STAND_IN and open('x.hdf5', 'rb').read(8) == b'\\x89HDF\\r\\n\\x1a\\n' and not os.path.exists('x.hdf5.npz')\n"
            "assert np.array_equal(np.array(h5py.File('x.hdf5', 'r')['Displacement']), np.eye(2))")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.parametrize("m,size,grid", [(4, 8, None), (3, 4, None), (5, 2, (1, 2, 1)), (4, 8, (2, 1, 4))])
def test_block_partition_of_the_structured_beam(m, size, grid):
    """device_setup.block_partition (host epart of the px x py x pz block grid): whole hexahedra per rank, rank =
    (bx*py + by)*pz + bz with block b owning hexahedra [b*n//p, (b+1)*n//p) along each axis, element order ascending."""
    nx, ny, nz = mesh.structured_beam_dims(m)
    px, py, pz = grid or ds.block_grid(size)
    assert px * py * pz == size
    ep = ds.block_partition(m, size, grid=grid)
    assert ep.shape == (6 * nx * ny * nz,) and set(np.unique(ep)) == set(range(size))
    hexes = ep.reshape(-1, 6)
    assert (hexes == hexes[:, :1]).all()                         # the six tetrahedra of a hexahedron stay together
    H = hexes[:, 0].reshape(nx, ny, nz)
    for r in range(size):
        bz, by, bx = r % pz, (r // pz) % py, r // (pz * py)
        sl = tuple(slice((b * n) // p, ((b + 1) * n) // p) for b, n, p in ((bx, nx, px), (by, ny, py), (bz, nz, pz)))
        assert (H[sl] == r).all() and (H == r).sum() == np.prod([s.stop - s.start for s in sl])
    # with P = 8 the partition has edges held by 4 ranks and the beam's centre line held by 8 (ascending-rank sums with >= 3 holders)
    if (px, py, pz) == (2, 2, 2):
        pts, cells, _ = mesh.structured_beam(m)
        holders = np.zeros((len(pts), size), dtype=bool)
        holders[cells.reshape(-1), np.repeat(ep, 4)] = True
        assert holders.sum(1).max() == 8 and (holders.sum(1) == 4).any()


def test_plan_cache_fingerprints_notice_changed_inputs():
    """Tools/Dynamic_solver.py caches the device plan per LocalK object; the cache signature covers F_rankwise, l_M and
    Local_Dirichlet (address, shape, sampled content), so replacing or rescaling one of them cannot leave stale copies."""
    pkg = os.path.join(ROOT, "synchronization-avoiding-algorithms_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from Tools.Dynamic_solver import _fingerprint
    F = np.arange(9000, dtype=np.float64).reshape(-1, 1)
    f0 = _fingerprint(F)
    assert _fingerprint(F) == f0
    assert _fingerprint(F.copy()) != f0                          # another array, even with equal content: new address
    F *= 2.0
    assert _fingerprint(F) != f0                                 # same array, rescaled in place
    D = list(range(0, 300, 3))
    assert _fingerprint(D) == _fingerprint(list(D)) and _fingerprint(D) != _fingerprint(D[:-1]) and _fingerprint(D) != _fingerprint([d + 1 for d in D])


def test_host_result_pool_policy(monkeypatch):
    """StepPlan._host_out (the array step_host returns d1 in): small plans get plain numpy memory; from 1 MiB per vector
    on, views of <= PINNED_POOL page-locked buffers, a buffer being reused only when no array — views and views of views
    included — refers to it any more; a caller that keeps every result falls back to plain memory beyond the pool.
    Page-locked memory needs the CUDA driver, so the allocation itself is replaced by ordinary memory here."""
    import ctypes

    class FakePinned:
        made = 0

        def __init__(self, n):
            FakePinned.made += 1
            self.buf = (ctypes.c_double * int(n))()
            self.__array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (ctypes.addressof(self.buf), False), "version": 3}

    monkeypatch.setattr(splan, "_PinnedVector", FakePinned)
    small = splan.StepPlan.__new__(splan.StepPlan)
    small.n_dof, small.h = 300, None
    a = small._host_out()
    assert a.shape == (300,) and a.base is None and FakePinned.made == 0
    pl = splan.StepPlan.__new__(splan.StepPlan)
    pl.n_dof, pl.h = 1 << 17, None
    # the reference's rotation: d_n, d_0 and the new d1 are alive at any time (+ the plan keeps the previous d0)
    d_n, d_0, seen = np.zeros(pl.n_dof), np.zeros(pl.n_dof), set()
    for i in range(20):
        d1 = pl._host_out()
        assert d1.flags.writeable and d1.dtype == np.float64 and d1.shape == (pl.n_dof,)
        assert d1.ctypes.data not in (d_0.ctypes.data, d_n.ctypes.data)         # never a buffer still in use
        d1[:] = i
        seen.add(d1.ctypes.data)
        pl._host_prev_d0 = d_0
        d_n, d_0 = d_0, d1.reshape(-1, 1)                                        # the shim returns (3n,1) views
        assert d_n.reshape(-1)[0] == max(i - 1, 0) and d_0[0, 0] == i
    assert len(seen) <= splan.PINNED_POOL and FakePinned.made == len(pl._pinned_pool) <= splan.PINNED_POOL
    # a slice of a view keeps its buffer busy
    tail = d_0[5:9]
    del d_0, d_n, d1
    pl._host_prev_d0 = None
    busy = tail.ctypes.data
    for _ in range(6):
        x = pl._host_out()
        assert not (x.ctypes.data <= busy < x.ctypes.data + 8 * pl.n_dof)
    # a caller that keeps everything: the pool is exhausted, plain memory from then on, nothing is overwritten
    keep = [pl._host_out() for _ in range(splan.PINNED_POOL + 3)]
    for i, k in enumerate(keep):
        k[:] = 100 + i
    assert all(k[0] == 100 + i for i, k in enumerate(keep)) and len({k.ctypes.data for k in keep}) == len(keep)
    assert sum(k.base is None for k in keep) >= 3 and len(pl._pinned_pool) == splan.PINNED_POOL
    monkeypatch.setenv("SAA_STEP_HOST_PINNED", "0")
    del keep
    assert pl._host_out().base is None


def test_shim_picks_the_gpu_of_the_local_rank(monkeypatch):
    """Tools/Dynamic_solver.py: the plan of a rank lives on GPU (local rank mod visible GPUs); SAA_DEVICE overrides."""
    pkg = os.path.join(ROOT, "synchronization-avoiding-algorithms_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    import Tools.Dynamic_solver as dsol
    monkeypatch.setattr(dsol._plan, "device_count", lambda: 4)
    for var in ("SAA_DEVICE", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID"):
        monkeypatch.delenv(var, raising=False)
    assert [dsol._device_for(r) for r in range(6)] == [0, 1, 2, 3, 0, 1]
    monkeypatch.setenv("OMPI_COMM_WORLD_LOCAL_RANK", "2")
    assert dsol._device_for(7) == 2
    monkeypatch.setenv("LOCAL_RANK", "5")
    assert dsol._device_for(7) == 1
    monkeypatch.setenv("SAA_DEVICE", "3")
    assert dsol._device_for(0) == 3
    monkeypatch.setattr(dsol._plan, "device_count", lambda: 0)
    assert dsol._device_for(9) == 0
