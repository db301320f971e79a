"""The drop-in `Tools` package (reference call surface) and the N>1 host logic.

CPU: set-up functions under their reference names vs the golden fixtures; world_size 2/3/4 gloo runs of the
host-side exchange logic.  GPU: the Data_prepare-shaped example driver, serial and with 2 processes, against
the golden displacement histories, bit for bit.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import ROOT, bits_equal, load_golden, read_result

PKG = os.path.join(ROOT, "synchronization-avoiding-algorithms_b200")


def _env():
    e = dict(os.environ)
    e["PYTHONPATH"] = os.pathsep.join([PKG, e.get("PYTHONPATH", ""), os.path.join(PKG, "compat")]).replace("::", ":")
    e.setdefault("OMP_NUM_THREADS", "1")
    return e


def _torchrun(n, script, *args, timeout=600, env_extra=None):
    port = 29500 + (os.getpid() * 7 + n) % 500
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), script, *args]
    r = subprocess.run(cmd, env=dict(_env(), **(env_extra or {})), capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def _tools():
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    import Tools.commons as c
    import Tools.Distributed_tools as d
    import Tools.Mat_construction as m
    import Tools.Steady_solvers as s
    return c, d, m, s


@pytest.mark.parametrize("name", ["beam_coarse_P2", "struct_m2_P4"])
def test_setup_through_reference_names(name):
    c, d, m, s = _tools()
    g = load_golden(name)
    Points, Cells = g["points"], g["cells"]
    E, nu, rho, fz = 1e6, 0.3, 1, 0.5
    elas = c.elasticity(E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu)), rho, fz, True)
    steady = c.elasticity(elas.lmd, elas.mu, rho, fz, False)
    assert bits_equal(elas.D(), np.array([[elas.lmd + 2.0 * elas.mu, elas.lmd, elas.lmd, 0, 0, 0],
                                          [elas.lmd, elas.lmd + 2.0 * elas.mu, elas.lmd, 0, 0, 0],
                                          [elas.lmd, elas.lmd, elas.lmd + 2.0 * elas.mu, 0, 0, 0],
                                          [0, 0, 0, elas.mu, 0, 0], [0, 0, 0, 0, elas.mu, 0], [0, 0, 0, 0, 0, elas.mu]], dtype=float))
    assert c.linear_ramp(0.3) == 0.3 and c.linear_ramp(1) == 1 and c.linear_ramp(2.5) == 1.0
    assert np.array_equal(c.node_to_dof(3, [0, 1, 2], [4, 0]), [12, 13, 14, 0, 1, 2])
    M0, _, F = m.Global_Assembly_no_bc(1, Cells, Points, steady, 0)
    lM = c.lumping_to_vec(M0)
    assert np.abs(lM - g["lumped_M"]).max() <= 4e-16 * np.abs(g["lumped_M"]).max()
    assert np.abs(F - g["F_pre"]).max() <= 4e-16 * np.abs(g["F_pre"]).max()
    assert np.abs(np.asarray(M0).sum(1) - lM[:, 0]).max() < 1e-15
    lists = []
    for r in range(g["P"]):
        ele, nodes = d.rankwise_dist(r, g["epart"], Points, Cells)
        lists.append(nodes)
        assert np.array_equal(ele, g["ranks"][r]["ele"]) and np.array_equal(nodes, g["ranks"][r]["nodes"])
        K = m.Local_assembly_for_stiffness(nodes, Cells[ele], Points, 1, 4, elas, r)
        assert np.array_equal(K.indices, g["ranks"][r]["K_indices"])
        assert np.abs(K.data - g["ranks"][r]["K_data"]).max() <= 1e-15 * np.abs(K.data).max()
        assert np.array_equal(d.Dirichlet_rank_dist(g["Dirichlet_node"], nodes), g["ranks"][r]["dirichlet"])
        dt = 0.9 * c.Meshsize(Cells[ele, :], Points) / np.sqrt(E / rho / (1 - nu ** 2))
        assert dt >= float(g["dt"])
    for r in range(g["P"]):
        sh = d.find_shared_nodes(r, g["P"], [len(x) for x in lists], lists)
        assert np.array_equal(sh, g["ranks"][r]["shared"])
        assert np.array_equal(c.node_to_dof(3, [0, 1, 2], d.local_mat_node(sh, lists[r])), g["ranks"][r]["loc_dof_shared"])
    # steady solution: equilibrium K d = F on the free DOFs, clamped DOFs zero
    Dd = c.node_to_dof(3, [0, 1, 2], g["Dirichlet_node"])
    dst = s.Steady_Elasticity_solver(1, Cells, Points, Dd, steady)
    Mm, Kk, Ff = m.Global_Assembly(1, Cells, Points, Dd, steady, t=None, steady=True)
    assert np.abs(dst[Dd]).max() == 0 and np.abs(Kk @ dst - Ff).max() < 1e-6 * np.abs(Ff).max()


def test_compat_h5py_and_meshio_standins(tmp_path):
    sys.path.append(os.path.join(PKG, "compat"))
    try:
        import h5py
        import meshio
        if not hasattr(meshio, "_mesh"):
            pytest.skip("real meshio installed")
        from saa_b200 import mesh
        p, c, f = mesh.structured_beam(1, length=2, with_facets="all")
        mesh.write_vtk(str(tmp_path / "a.vtk"), p, c, f)
        M = meshio.read(str(tmp_path / "a.vtk"))
        assert bits_equal(M.points, p) and np.array_equal(M.cells_dict["tetra"], c) and np.array_equal(M.cells_dict["triangle"], f)
        meshio.write_points_cells(str(tmp_path / "b.vtk"), M.points, M.cells, {"u": np.arange(len(p), dtype=float)})
        assert np.array_equal(meshio.read(str(tmp_path / "b.vtk")).cells_dict["tetra"], c)
        a = np.arange(12.0).reshape(3, 4)
        hf = h5py.File(str(tmp_path / "x.hdf5"), "w")
        hf.create_dataset("Displacement", data=a, compression="gzip")
        hf.close()
        assert np.array_equal(h5py.File(str(tmp_path / "x.hdf5"), "r")["Displacement"][:], a)
        assert open(str(tmp_path / "x.hdf5"), "rb").read(4) == b"\x89HDF"          # genuine HDF5, stand-in or not
    finally:
        sys.path.remove(os.path.join(PKG, "compat"))


@pytest.mark.parametrize("name,n", [("beam_coarse_P2", 2), ("beam_coarse_P3", 3), ("struct_m2_P4", 4)])
def test_multirank_host_logic_gloo(name, n):
    """world_size n on CPU (gloo): facade collectives, distributed partition call, maps, neighbour exchange +
    ascending-rank sum == literal syn_cpus."""
    out = _torchrun(n, os.path.join(ROOT, "tests", "dist_worker.py"), "host", name)
    assert out.count("ok (host)") == n


# ---------------------------------------------------------------------------------------------------------
def _run_driver(tmp_path, g, nproc, steps, env_extra=None):
    from saa_b200 import mesh
    vtk = str(tmp_path / "mesh.vtk")
    mesh.write_vtk(vtk, g["points"], g["cells"], g["facets"])
    script = os.path.join(ROOT, "examples", "data_prepare_driver.py")
    args = ["--mesh", vtk, "--steps", str(steps), "--out", str(tmp_path), "--steady"]
    log = ""
    if nproc == 1:
        r = subprocess.run([sys.executable, script, *args], env=_env(), capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    else:
        log = _torchrun(nproc, script, *args, timeout=900, env_extra=env_extra)
    _run_driver.log = log
    out = []
    for q in range(nproc):
        out.append(read_result(tmp_path / "Results" / "Dynamics" / f"Local-rank-{q}.hdf5"))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name,steps", [("beam_coarse_P1", 2000), ("beam_coarse_P2", 1000)])
def test_data_prepare_shaped_driver_matches_reference_history(tmp_path, name, steps):
    """The reference workflow (config 1: beam_coarse, np = 1 and np = 2) through `Tools.*` names; two processes
    share the one GPU of the box and exchange through gloo host messages."""
    g = load_golden(name)
    H = _run_driver(tmp_path, g, g["P"], steps)
    for q in range(g["P"]):
        assert H[q].shape == (g["ranks"][q]["F"].size, steps)
        for n in [int(s) for s in g["steps"] if s <= steps]:
            assert bits_equal(H[q][:, n - 1], g[f"hist_{n}_r{q}"]), (name, n, q)
    sh = np.loadtxt(str(tmp_path / "Results" / "Shared_Data" / "Rank=0_shared.csv"), dtype=np.int64, ndmin=1)
    assert np.array_equal(sh, g["ranks"][0]["shared"])
    assert os.path.isfile(str(tmp_path / "Results" / "Static" / "steady_distributed.vtk"))


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["auto", "host"])
def test_shim_halo_transport_two_processes(tmp_path, transport):
    """parallel_explicit_solver_dis_pre with size = 2 under torchrun: by default the shim maps the neighbours' receive
    areas (CUDA IPC; NVLink stores between GPUs, plain stores when the ranks share a GPU) and every MODEL=False call is one
    pipelined saa_step_host_ex(MODE_SYNC); SAA_SHIM_TRANSPORT=host keeps the messages on the caller's communicator.  Both
    reproduce the reference's history bit for bit."""
    g = load_golden("beam_coarse_P2")
    H = _run_driver(tmp_path, g, 2, 200, env_extra={"SAA_SHIM_TRANSPORT": transport, "SAA_SHIM_VERBOSE": "1"})
    assert _run_driver.log.count("halo transport: " + ("peer" if transport == "auto" else "host")) == 2, _run_driver.log[-2000:]
    for q in range(2):
        for n in (1, 2, 10, 100):
            assert bits_equal(H[q][:, n - 1], g[f"hist_{n}_r{q}"]), (transport, n, q)


@pytest.mark.gpu
def test_syn_cpus_on_device_two_processes():
    out = _torchrun(2, os.path.join(ROOT, "tests", "dist_worker.py"), "gpu", "beam_coarse_P2")
    assert out.count("ok (gpu)") == 2
