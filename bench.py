#!/usr/bin/env python
"""bench.py — DOF-steps/s of the explicit FE time step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--refine M] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE explicit time step (one pass of Dynamic_solver.py:9-34 over the whole mesh): force K.u,
central-difference update, Dirichlet clamp and — for N > 1 — the shared-node force exchange.
Workload (config.workload): the 25 x 1 x 1 cantilever of Mesh_info/beam_US.geo as a structured tet mesh with
m cells per unit length.  Defaults follow BASELINE.json's configs: one GPU -> m = 24 (1.13 M DOF, "~1M DOF,
single B200, fp64"); N > 1 -> m = 65 (21.2 M DOF, "~20M DOF over 2/4/8 B200") cut into N x-slabs, one
process per GPU, shared-node forces exchanged every step.  --refine 111 is the 104 M-DOF mesh of the 100M sweep.

`value`   whole-job DOF-steps/s with the state resident in HBM (CUDA events on the plan's stream, max over ranks).
`e2e`     the same metric through the reference-facing call saa_step_host — one parallel_explicit_solver_dis_pre
          evaluation per call with (d0, dn) in pinned HOST memory and d1 returned to HOST memory every step.
`roofline` algorithmic bytes of the fused force+update kernel in the format it streams (76 B per stored 3x3 node
          block + five fp64 vector streams; the scalar-CSR figure of SURVEY.md §8d is reported beside it) / average
          step time, vs the measured HBM copy bandwidth.
`cpu_baseline` the CPU oracle (oracle/fem_oracle.c, OpenMP) on the same problem on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed time steps (0: 10000 for m <= 32, else 2000)")
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--refine", "--m", dest="m", type=int, default=0,
                    help="cells per unit length of the 25x1x1 beam (0: default for N); spell it --refine under torchrun")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed host-call steps (0: 200 for m <= 32, else 30)")
    ap.add_argument("--launch", default="auto", choices=["auto", "per_step", "graph", "persistent"])
    ap.add_argument("--setup", default="device", choices=["device", "host"], help="where the problem is assembled")
    ap.add_argument("--cpu-seconds", type=float, default=0.0, help="CPU time budget of the oracle runs (0: 15 s native arm, 20 s reference arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="N = 1: skip the extra 21 M-DOF measurement")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"], help="halo transport for N > 1")
    ap.add_argument("--sync-avoid", default="", help="N > 1: also time the synchronization-avoiding loop (BASELINE config 5); "
                    "comma list of re-sync periods k, 0 = never re-synchronise (the reference's behaviour), e.g. 0,10,50")
    ap.add_argument("--balance", action="store_true", help="N > 1: size the x-slabs by each GPU's measured speed (a short local "
                    "run first), so that a slower GPU does not pace the others")
    ap.add_argument("--filter-size", type=int, default=150, help="n_s of the LSTM refill (Online_predictor.py:57)")
    return ap.parse_args()


def default_m(n_gpus):
    """BASELINE.json configs: one B200 -> ~1M DOF (m = 24: 1 126 875 DOF); 2/4/8 B200 -> ~20M DOF (m = 65:
    21 215 700 DOF) partitioned over the GPUs.  --m overrides (m = 111: 104 M DOF, the 100M-DOF sweep)."""
    return 24 if n_gpus == 1 else 65


def workload_name(m, n_dof, n_el):
    cfg = {24: "BASELINE config 2 (~1M DOF, single B200, fp64)", 65: "BASELINE config 3 (~20M DOF over 2/4/8 B200)",
           111: "BASELINE config 4 (~100M DOF strong-scaling sweep)"}.get(m, "custom refinement")
    return f"structured 25x1x1 cantilever (Mesh_info/beam_US.geo box) m={m}: {n_dof} DOF, {n_el} tets; {cfg}"


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
            # median over the samples taken under load (power above half of the maximum seen)
            busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


# ---- problem set-up -------------------------------------------------------------------------------------------
def setup_host(m, size, rank, local, make_plan=True):
    """Host set-up (bit-exact with the reference's own assembly): numpy/scipy, METIS partition for N > 1."""
    import saa_b200  # noqa: F401
    from saa_b200 import mesh, partition, problem
    pts, cells, fac = mesh.structured_beam(m)
    if size == 1:
        ep, part = np.zeros(len(cells), dtype=np.int64), "none"
    else:
        ep, part = partition.metis_part_mesh(cells, len(pts), size), "METIS_PartMeshDual(ncommon=3)"
    pb = problem.build_problem(pts, cells, fac, ep, size, ranks=[rank])
    q = pb["ranks"][rank]
    pl = problem.make_plan(q, pb["dt"], problem.DAMP_DEFAULT, size, device=local) if make_plan else None
    csr = dict(K=q["K"], F=q["F"], lM=q["lM"], dirichlet=q["dirichlet"], nodes=q["nodes"])
    return pl, dict(dt=float(pb["dt"]), n_nodes=len(pts), n_elem=len(cells), part=part, csr=csr,
                    assembly="host (numpy/scipy, bit-exact with the reference's assembly)")


def setup_device(m, size, rank, local, keep_csr=False):
    """Device set-up (saa_b200.device_setup): slab mesh, numbering, K6 assembly and the plan, all on the GPU."""
    import saa_b200  # noqa: F401
    from saa_b200 import device_setup
    pl, info = device_setup.build_structured_rank(m, rank, size, device_index=local, keep_csr=keep_csr)
    csr = None
    if keep_csr:
        csr = dict(K=info["K"].to_scipy(), F=info["F"].cpu().numpy(), lM=info["lM"].cpu().numpy(), dirichlet=info["dirichlet"],
                   nodes=info["local_nodes"].cpu().numpy())
        info["K"].free()
    return pl, dict(dt=float(info["dt"]), n_nodes=info["n_global_nodes"], n_elem=info["n_global_elem"],
                    part="none" if size == 1 else f"{size} x-slabs of whole hexahedron layers (what a k-way cut of a 25:1:1 beam gives)",
                    csr=csr, assembly="device (saa_assemble_stiffness_dev, closed-form element matrices)")


def oracle_for(csr, n_nodes, dt):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fem_oracle
    K = csr["K"]
    ranks = [dict(K_indptr=K.indptr, K_indices=K.indices, K_data=K.data, F=csr["F"], lM=csr["lM"], dirichlet=csr["dirichlet"],
                  nodes=csr["nodes"])]
    return fem_oracle.OracleProblem(n_nodes, ranks, dt, 0.5)


def cpu_baseline(csr, n_nodes, dt, seconds):
    """Time the C oracle (all OpenMP threads) on the SAME serial problem for a bounded number of steps."""
    o = oracle_for(csr, n_nodes, dt)
    n_dof = csr["F"].size
    o.run(2)                                  # warm
    t0 = time.perf_counter(); o.run(1); t1 = time.perf_counter() - t0
    k = int(max(3, min(2000, seconds / max(t1, 1e-6))))
    t0 = time.perf_counter(); o.run(k); secs = time.perf_counter() - t0
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threads = int(os.environ.get("OMP_NUM_THREADS", cores))
    o.close()
    return dict(value=n_dof * k / secs, unit="DOF-steps/s", cores=threads, kind="port",
                sample=f"{k} consecutive time steps of the same {n_dof}-DOF mesh with oracle/fem_oracle.c (OpenMP over rows, "
                       f"{threads} threads); the reference itself is single-threaded numpy/scipy per MPI rank"), k, secs


def run_reference(args, emit):
    """--impl reference: the CPU restatement of the reference's own path on the host cores (the reference is
    Python and does not exist on the GPU box; oracle/fem_oracle.c is its pinned, bit-exact port).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    os.environ.pop("OMP_NUM_THREADS", None)   # torchrun pins it to 1; the baseline may use every core
    m = args.m or default_m(args.gpus)
    csr = None
    if m > 32:
        try:                                   # inputs of the CPU run come from the device assembly (host one would take minutes)
            pl, info = setup_device(m, 1, 0, 0, keep_csr=True)
            csr, n_nodes, n_el, dt = info["csr"], info["n_nodes"], info["n_elem"], info["dt"]
            pl.close()
        except Exception as e:                 # no GPU: fall through to the host assembly
            print(f"[bench] device set-up unavailable for the reference arm ({e}); assembling on the host", file=sys.stderr)
    if csr is None:
        _, info = setup_host(m, 1, 0, 0, make_plan=False)
        csr, n_nodes, n_el, dt = info["csr"], info["n_nodes"], info["n_elem"], info["dt"]
    cb, k, secs = cpu_baseline(csr, n_nodes, dt, args.cpu_seconds or 20.0)
    n_dof = 3 * n_nodes
    line = {"impl": "reference", "metric": "DOF-steps/sec", "value": cb["value"], "unit": "DOF-steps/s",
            "n_gpus": args.gpus, "steps": k, "warmup": 3, "ms_per_step": 1e3 * secs / k, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(m, n_dof, n_el),
                       "note": f"requested --steps {args.steps or 'default'} bounded to {k} CPU time steps of the full mesh"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "DOF-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def time_resident(pl, torch, stream, steps, warmup, mode, launch, barrier):
    pl.step(warmup, mode, launch)
    pl.synchronize()
    barrier()
    l0 = pl.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    pl.step(steps, mode, launch)
    e1.record(stream)
    pl.synchronize()
    barrier()
    return e0.elapsed_time(e1), pl.kernel_launches - l0


def time_sync_avoiding(pl, args, torch, stream, barrier, max_over_ranks, n_dof_global, local):
    """BASELINE config 5: the loop of Online_predictor.py:251-318 on the device — LSTM encoder-decoder (the reference's
    architecture: 2-layer bidirectional encoder, hidden 50, n_past = n_future = 20; random-init weights, synthetic)
    predicts this rank's shared-DOF displacements on the GPU, un-synchronised steps overwrite them; a true
    exchange runs every k steps (k = 0: never again after the warm-up, the reference's behaviour)."""
    from saa_b200 import sync_avoiding
    sys.path.insert(0, os.path.join(ROOT, "synchronization-avoiding-algorithms_b200"))
    from Tools.DNN_tools import LSTM_encoder_decoder
    nb, off = pl.halo_layout()
    hp = pl._halo_desc
    dofs = (3 * np.asarray(hp["shared_pos"], dtype=np.int64)[:, None] + np.arange(3)[None, :]).ravel()
    n_p = n_f = 20
    n_s = args.filter_size
    out = {"n_past": n_p, "n_future": n_f, "filter_size": n_s, "input_size_rank0": int(dofs.size), "hidden": 50,
           "model": "LSTM_encoder_decoder(input, 50, 2, bidirectional) random-init, fp32, on-device (PyTorch/cuDNN)", "runs": []}
    torch.manual_seed(1234 + pl.rank)
    model = LSTM_encoder_decoder(int(dofs.size), 50, 2, True, 0.0, 0.0)
    # one untimed inference first: cuDNN initialises / picks its LSTM algorithm on the first call (~0.3 s)
    model.to(f"cuda:{local}").eval()
    sync_avoiding._dnn_prediction().predict_block(
        model, torch.zeros((n_p * n_s, int(dofs.size)), dtype=torch.float64, device=f"cuda:{local}"), n_p, n_f, n_s, 1e-3, -1e-2)
    torch.cuda.synchronize()
    for k in [int(x) for x in args.sync_avoid.split(",")]:
        run = sync_avoiding.SyncAvoidingRun([pl], pl, [dofs], [model], [(1e-3, -1e-2)], n_p, n_f, n_s, device=f"cuda:{local}",
                                            resync_every=(k or None))
        # the history ring starts at this call: warm-up = n_p*n_s synchronised steps, then whole refill blocks are timed
        run.i = 0
        base = pl.history_count
        run.run(n_p * n_s)
        pl.synchronize()
        barrier()
        blocks = max(1, min(3, 6000 // (n_f * n_s)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        run.run(n_p * n_s + blocks * n_f * n_s)
        pl.synchronize()
        barrier()
        ms = max_over_ranks((time.perf_counter() - w0) * 1e3)         # wall clock: includes the LSTM inference of every block
        nst = blocks * n_f * n_s
        out["runs"].append({"resync_every": k, "steps": nst, "ms_per_step": ms / nst, "value": n_dof_global * nst / (ms * 1e-3),
                            "unit": "DOF-steps/s", "lstm_ms_per_block_rank0": 1e3 * run.t_predict / max(1, blocks + 0)})
        pl.set_history(None, 0, 1)
    return out


def main():
    args = parse()
    # keep stdout clean for the single JSON line: libraries (NCCL's version banner, torchrun notices) write to it
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)

    if args.impl == "reference":
        run_reference(args, emit)
        return
    import torch
    import torch.distributed as dist
    import saa_b200  # noqa: F401
    from saa_b200 import multi, plan as splan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the time-step path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    m = args.m or default_m(world)
    steps = args.steps or (10000 if m <= 32 else 2000)
    e2e_steps = args.e2e_steps or (200 if m <= 32 else 30)
    want_cpu = (not args.no_cpu_baseline) and world == 1
    t_setup = time.time()
    balance = None
    if args.setup == "host":
        pl, info = setup_host(m, world, rank, local)
    else:
        pl, info = setup_device(m, world, rank, local, keep_csr=want_cpu and m <= 32)
        if world > 1 and args.balance:
            # pass 1 built equal slabs: time them without any exchange, then rebuild with layers ~ measured speed
            from saa_b200 import device_setup
            st0 = torch.cuda.ExternalStream(pl.stream, device=torch.device("cuda", local))
            ms0, _ = time_resident(pl, torch, st0, 300, 30, splan.MODE_LOCAL, splan.LAUNCH_AUTO, lambda: torch.cuda.synchronize())
            speed = torch.tensor([pl.n_dof / ms0], dtype=torch.float64, device="cuda")
            allv = [torch.zeros_like(speed) for _ in range(world)]
            dist.all_gather(allv, speed)
            balance = [float(v.item()) for v in allv]
            pl.close()
            del pl
            torch.cuda.empty_cache()
            device_setup.set_layer_weights(balance)
            pl, info = setup_device(m, world, rank, local)
            balance = [b / max(balance) for b in balance]
    transport = multi.attach_transport(pl, args.transport) if world > 1 else "none"
    t_setup = time.time() - t_setup
    n_dof_global = 3 * info["n_nodes"]
    n_dof_local = pl.n_dof
    dtv = info["dt"]
    mode = splan.MODE_SYNC if world > 1 else splan.MODE_LOCAL
    launch = {"auto": splan.LAUNCH_AUTO, "per_step": splan.LAUNCH_PER_STEP, "graph": splan.LAUNCH_GRAPH,
              "persistent": splan.LAUNCH_PERSISTENT}[args.launch]
    stream = torch.cuda.ExternalStream(pl.stream, device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing --------------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches = time_resident(pl, torch, stream, steps, args.warmup, mode, launch, barrier)
    ms = max_over_ranks(ms)
    ms_local = None
    if world > 1:   # the same shards stepped WITHOUT the exchange (MODEL=True arithmetic): what the halo costs per step
        k_loc = min(steps, 500)
        ms_local, _ = time_resident(pl, torch, stream, k_loc, 20, splan.MODE_LOCAL, launch, barrier)
        ms_local = max_over_ranks(ms_local) / k_loc

    # ---- end to end through the reference-facing host call -----------------------------------------
    d0, dn, tn = pl.get_state()
    h0 = torch.from_numpy(d0).pin_memory().numpy()
    hn = torch.from_numpy(dn).pin_memory().numpy()
    h1 = torch.empty(n_dof_local, dtype=torch.float64).pin_memory().numpy()
    for _ in range(3):
        pl.step_host(h0, hn, tn, mode, out=h1)
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        pl.step_host(h0, hn, tn, mode, out=h1)       # d1 lands in host memory every call
        h0, hn, h1 = h1, h0, hn                      # d_n = d_0; d_0 = d1 (Data_prepare.py:233-234)
        tn = tn + dtv
    pl.synchronize()
    barrier()
    ms_e2e = max_over_ranks((time.perf_counter() - w0) * 1e3)   # the call is synchronous: wall clock covers copies + kernels
    clocks = sampler.stop() if sampler else None

    # ---- roofline of the fused force+update kernel (largest shard bounds the step) ------------------
    # algorithmic bytes of the format actually streamed (node-block sliced ELL): 76 B per stored 3x3 block (nine
    # fp64 values + one int32 column-node id) + slice offsets + Dirichlet mask words + five fp64 vector streams.
    # The scalar-CSR figure of SURVEY.md §8d (12 B per stored entry + 4 B per row pointer + 40 B per row) is given
    # beside it as csr_equivalent.
    nnz = pl.nnz
    alg_bytes = max_over_ranks(float(pl.matrix_bytes + 5 * 8 * n_dof_local))
    csr_bytes = max_over_ranks(float(nnz * 12 + (n_dof_local + 1) * 4 + 5 * 8 * n_dof_local))
    step_s = ms * 1e-3 / steps
    peak, peak_src = measured_peak()
    achieved = alg_bytes / step_s / 1e9
    traffic = None
    tf = os.path.join(ROOT, "profiles", "traffic_r1.json")
    if os.path.isfile(tf) and m == 24 and world == 1:
        try:
            traffic = json.load(open(tf)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    also = None
    if world == 1 and m == 24 and not args.no_also and args.setup == "device":
        # the strong-scaling base: BASELINE config 3's 21 M-DOF mesh on this one GPU
        try:
            pl2, info2 = setup_device(65, 1, 0, local)
            st2 = torch.cuda.ExternalStream(pl2.stream, device=torch.device("cuda", local))
            ms2, _ = time_resident(pl2, torch, st2, 1000, 50, splan.MODE_LOCAL, launch, barrier)
            b2 = pl2.matrix_bytes + 40 * pl2.n_dof
            also = {"workload": workload_name(65, 3 * info2["n_nodes"], info2["n_elem"]), "steps": 1000,
                    "value": 3 * info2["n_nodes"] * 1000 / (ms2 * 1e-3), "ms_per_step": ms2 / 1000,
                    "roofline_frac": b2 / (ms2 * 1e-6) / 1e9 / peak, "nnz_per_row": pl2.nnz / pl2.n_dof}
            pl2.close()
        except Exception as e:
            also = {"error": str(e)[:200]}

    sync_avoid = None
    if world > 1 and args.sync_avoid:
        sync_avoid = time_sync_avoiding(pl, args, torch, stream, barrier, max_over_ranks, n_dof_global, local)

    if rank == 0:
        line = {
            "metric": "DOF-steps/sec", "value": n_dof_global * steps / (ms * 1e-3), "unit": "DOF-steps/s",
            "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(m, n_dof_global, info["n_elem"]),
                       "partition": info["part"], "balance": balance, "transport": transport, "assembly": info["assembly"],
                       "nnz_per_row": nnz / n_dof_local, "local_dof_rank0": n_dof_local,
                       "launch": args.launch, "dt": dtv, "setup_s": round(t_setup, 1), "ms_per_step_without_exchange": ms_local,
                       "l2": "inputs larger than L2: matrix stream per step per GPU = %.0f MB vs 126 MB L2" % (pl.matrix_bytes / 1e6)},
            "e2e": {"value": n_dof_global * e2e_steps / (ms_e2e * 1e-3), "unit": "DOF-steps/s",
                    "h2d_bytes_per_step": 2 * 8 * n_dof_local, "d2h_bytes_per_step": 8 * n_dof_local,
                    "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                    "call": "saa_step_host (one parallel_explicit_solver_dis_pre evaluation per call, pinned host d0/dn in, d1 out)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_launch": alg_bytes, "bytes_formula": "76*blocks + 8*(slices+1) + 4*rows/32 + 40*rows (largest shard)",
                         "csr_equivalent": {"bytes_per_launch": csr_bytes, "formula": "12*nnz + 4*(rows+1) + 40*rows",
                                            "achieved": csr_bytes / step_s / 1e9, "frac": csr_bytes / step_s / 1e9 / peak},
                         "blocks_per_node": (pl.padded_entries / 9) / (n_dof_local / 3),
                         "kernel": "saa_k_step (fused K.u + central-difference update + Dirichlet mask)"},
            "clocks": clocks,
        }
        if also is not None:
            line["also"] = also
        if sync_avoid is not None:
            line["sync_avoiding"] = sync_avoid
        if want_cpu:
            csr = info["csr"]
            if csr is None:     # large mesh: time the CPU on the 1.13 M-DOF mesh instead (DOF-normalised metric)
                _, i24 = setup_host(24, 1, 0, local, make_plan=False)
                csr, nn, dt24 = i24["csr"], i24["n_nodes"], i24["dt"]
            else:
                nn, dt24 = info["n_nodes"], dtv
            cb, _, _ = cpu_baseline(csr, nn, dt24, args.cpu_seconds or 15.0)
            line["cpu_baseline"] = cb
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
