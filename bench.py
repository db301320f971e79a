#!/usr/bin/env python
"""bench.py — DOF-steps/s of the explicit FE time step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--m M] [--impl native|reference]

A "step" is ONE explicit time step (one pass of Dynamic_solver.py:9-34 over the whole mesh):
force K.u, central-difference update, Dirichlet clamp and — for N > 1 — the shared-node force exchange.
Workload (config.workload): the 25 x 1 x 1 cantilever of Mesh_info/beam_US.geo as a structured tet mesh
with m cells per unit length; m = 24 -> 1.13 M DOF (BASELINE config 2, "~1M DOF, single B200, fp64").
For N > 1 the same mesh is partitioned over the N GPUs (one process per GPU, halo exchange over NCCL).

`value`  : whole-job DOF-steps/s with the state resident in HBM (CUDA events on the plan's stream).
`e2e`    : the same metric through the reference-facing call saa_step_host — one
           parallel_explicit_solver_dis_pre evaluation per call with (d0, dn) in pinned HOST memory
           and d1 returned to HOST memory, copies inside the timed region.
`roofline`: algorithmic bytes of the fused force+update kernel (12 B per stored entry
           + 4 B per row pointer + five fp64 vector streams) / average step time, vs the measured
           HBM copy bandwidth of MEASURED_PEAKS.json.
`cpu_baseline`: the CPU oracle (oracle/fem_oracle.c, OpenMP) on the same problem on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--m", type=int, default=0, help="cells per unit length of the 25x1x1 beam (0: default for N)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--launch", default="auto", choices=["auto", "per_step", "graph", "persistent"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"], help="halo transport for N > 1")
    return ap.parse_args()


def default_m(n_gpus):
    return 24  # 1.13 M DOF (config 2); the same mesh is strong-scaled over N GPUs


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            t = [x.strip() for x in line.split(",")]
            if len(t) < 7:
                continue
            try:
                sm.append(float(t[0])); mx.append(float(t[1])); pw.append(float(t[2]))
            except ValueError:
                continue
            for nm, v in zip(names, t[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


def build_rank_problem(m, size, rank):
    """Mesh, partition and this rank's assembled inputs (host side; see saa_b200.problem)."""
    import pickle
    import saa_b200  # noqa: F401
    from saa_b200 import mesh, partition, problem
    cache = os.environ.get("SAA_BENCH_CACHE")       # optional: reuse the assembled problem across runs of one session
    cfile = os.path.join(cache, f"pb_m{m}_P{size}_r{rank}.pkl") if cache else None
    if cfile and os.path.isfile(cfile):
        with open(cfile, "rb") as fh:
            return pickle.load(fh)
    pts, cells, fac = mesh.structured_beam(m)
    if size == 1:
        ep = np.zeros(len(cells), dtype=np.int64)
        part = "none"
    elif len(cells) <= 3_000_000:
        ep = partition.metis_part_mesh(cells, len(pts), size)
        part = "METIS_PartMeshDual(ncommon=3)"
    else:
        ep = partition.slab_partition(pts, cells, size)
        part = "x-slabs"
    pb = problem.build_problem(pts, cells, fac, ep, size, ranks=[rank])
    out = (pb, part, len(pts), len(cells))
    if cfile:
        os.makedirs(cache, exist_ok=True)
        with open(cfile, "wb") as fh:
            pickle.dump(out, fh, protocol=4)
    return out


def oracle_for(pb, rank_ids, n_nodes):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fem_oracle
    ranks = [dict(K_indptr=pb["ranks"][r]["K"].indptr, K_indices=pb["ranks"][r]["K"].indices,
                  K_data=pb["ranks"][r]["K"].data, F=pb["ranks"][r]["F"], lM=pb["ranks"][r]["lM"],
                  dirichlet=pb["ranks"][r]["dirichlet"], nodes=pb["ranks"][r]["nodes"]) for r in rank_ids]
    return fem_oracle.OracleProblem(n_nodes, ranks, pb["dt"], 0.5)


def cpu_baseline(pb, n_nodes, seconds, start_state=None):
    """Time the C oracle (all OpenMP threads) on the SAME serial problem for a bounded number of steps."""
    o = oracle_for(pb, [0], n_nodes)
    if start_state is not None:
        o.set_state(0, *start_state)
    n_dof = pb["ranks"][0]["F"].size
    o.run(2)                                  # warm
    t0 = time.perf_counter(); o.run(1); t1 = time.perf_counter() - t0
    k = int(max(3, min(2000, seconds / max(t1, 1e-6))))
    t0 = time.perf_counter(); o.run(k); dt_ = time.perf_counter() - t0
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threads = int(os.environ.get("OMP_NUM_THREADS", cores))
    o.close()
    return dict(value=n_dof * k / dt_, unit="DOF-steps/s", cores=threads, kind="port",
                sample=f"{k} consecutive time steps of the same {n_dof}-DOF mesh with oracle/fem_oracle.c (OpenMP over rows, "
                       f"{threads} threads); the reference itself is single-threaded numpy/scipy per MPI rank"), k, dt_


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's own path on the host cores (the reference is
    Python and is not present on the GPU box; oracle/fem_oracle.c is its pinned, bit-exact port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m = args.m or default_m(args.gpus)
    pb, part, n_nodes, n_el = build_rank_problem(m, 1, 0)
    cb, k, secs = cpu_baseline(pb, n_nodes, max(args.cpu_seconds, 20.0))
    n_dof = 3 * n_nodes
    line = {"impl": "reference", "metric": "DOF-steps/sec", "value": cb["value"], "unit": "DOF-steps/s",
            "n_gpus": args.gpus, "steps": k, "warmup": 3, "ms_per_step": 1e3 * secs / k, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"structured 25x1x1 cantilever m={m}: {n_dof} DOF, {n_el} tets (BASELINE config 2)",
                       "note": f"requested --steps {args.steps} bounded to {k} CPU time steps of the full mesh"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "DOF-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    import saa_b200  # noqa: F401
    from saa_b200 import multi, plan as splan, problem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the time-step path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    m = args.m or default_m(world)
    t_setup = time.time()
    pb, part, n_nodes, n_el = build_rank_problem(m, world, rank)
    q = pb["ranks"][rank]
    pl = problem.make_plan(q, pb["dt"], problem.DAMP_DEFAULT, world, device=local)
    transport = multi.attach_transport(pl, args.transport) if world > 1 else "none"
    t_setup = time.time() - t_setup
    n_dof_global = 3 * n_nodes
    n_dof_local = pl.n_dof
    mode = splan.MODE_SYNC if world > 1 else splan.MODE_LOCAL
    launch = {"auto": splan.LAUNCH_AUTO, "per_step": splan.LAUNCH_PER_STEP, "graph": splan.LAUNCH_GRAPH,
              "persistent": splan.LAUNCH_PERSISTENT}[args.launch]
    stream = torch.cuda.ExternalStream(pl.stream, device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing --------------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    pl.step(args.warmup, mode, launch)
    pl.synchronize()
    barrier()
    l0 = pl.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    pl.step(args.steps, mode, launch)
    e1.record(stream)
    pl.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = pl.kernel_launches - l0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # ---- end to end through the reference-facing host call -----------------------------------------
    d0, dn, tn = pl.get_state()
    h0 = torch.from_numpy(d0).pin_memory().numpy()
    hn = torch.from_numpy(dn).pin_memory().numpy()
    h1 = torch.empty(n_dof_local, dtype=torch.float64).pin_memory().numpy()
    dtv = float(pb["dt"])
    for _ in range(3):
        pl.step_host(h0, hn, tn, mode, out=h1)
    barrier()
    w0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.e2e_steps):
        pl.step_host(h0, hn, tn, mode, out=h1)       # d1 lands in host memory every call
        h0, hn, h1 = h1, h0, hn                      # d_n = d_0; d_0 = d1 (Data_prepare.py:233-234)
        tn = tn + dtv
    e1.record(stream)
    pl.synchronize()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    wall_e2e = (time.perf_counter() - w0) * 1e3
    t = torch.tensor([max(ms_e2e, wall_e2e)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    clocks = sampler.stop() if sampler else None

    # ---- roofline of the fused force+update kernel (local sizes of this rank; max time over ranks) ---
    nnz = pl.nnz
    alg_bytes = nnz * 12 + (n_dof_local + 1) * 4 + 5 * 8 * n_dof_local
    sums = torch.tensor([float(alg_bytes), float(nnz), float(n_dof_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.MAX)       # the slowest rank bounds the step: use the largest shard
    alg_bytes_max = float(sums[0].item())
    step_s = ms * 1e-3 / args.steps
    peak, peak_src = measured_peak()
    achieved = alg_bytes_max / step_s / 1e9
    traffic = None
    tf = os.path.join(ROOT, "profiles", "traffic_r1.json")
    if os.path.isfile(tf):
        try:
            traffic = json.load(open(tf)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    if rank == 0:
        line = {
            "metric": "DOF-steps/sec", "value": n_dof_global * args.steps / (ms * 1e-3), "unit": "DOF-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"structured 25x1x1 cantilever (Mesh_info/beam_US.geo box) m={m}: {n_dof_global} DOF, "
                                   f"{n_el} tets; BASELINE config 2 (~1M DOF, fp64)",
                       "partition": part, "transport": transport, "nnz_per_row": nnz / n_dof_local, "local_dof_rank0": n_dof_local,
                       "launch": args.launch, "dt": dtv, "setup_s": round(t_setup, 1),
                       "l2": "inputs larger than L2: matrix stream per step = %.0f MB vs 126 MB L2" % (pl.matrix_bytes / 1e6)},
            "e2e": {"value": n_dof_global * args.e2e_steps / (ms_e2e * 1e-3), "unit": "DOF-steps/s",
                    "h2d_bytes_per_step": 2 * 8 * n_dof_local, "d2h_bytes_per_step": 8 * n_dof_local,
                    "steps": args.e2e_steps, "ms_per_step": ms_e2e / args.e2e_steps,
                    "call": "saa_step_host (one parallel_explicit_solver_dis_pre evaluation per call, pinned host d0/dn in, d1 out)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_launch": alg_bytes_max, "bytes_formula": "12*nnz + 4*(rows+1) + 40*rows",
                         "stored_bytes_per_launch": pl.matrix_bytes + 40 * n_dof_local,
                         "kernel": "saa_k_step (fused K.u + central-difference update + Dirichlet mask)"},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            cb, _, _ = cpu_baseline(pb, n_nodes, args.cpu_seconds)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
