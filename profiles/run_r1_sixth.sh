#!/bin/bash
# sixth GPU pass (8 GPUs): P=4 / P=8 transports parity; strong scaling at N=4, 8 on the 21M- and 104M-DOF meshes
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
python -m pytest tests/test_gpu_parity.py -x -q -k "one_process_per_gpu" > gpurun_out/pytest_peer8.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_peer8.log
tail -5 gpurun_out/pytest_peer8.log
for N in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
  $TR bench.py --gpus $N --steps 4000 > gpurun_out/bench_n${N}_m65.json 2> gpurun_out/bench_n${N}_m65.err
  $TR bench.py --gpus $N --refine 111 --steps 2000 > gpurun_out/bench_n${N}_m111.json 2> gpurun_out/bench_n${N}_m111.err
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 4000 --transport nccl > gpurun_out/bench_n8_m65_nccl.json 2> gpurun_out/bench_n8_m65_nccl.err
$TR bench.py --gpus 8 --refine 24 > gpurun_out/bench_n8_m24.json 2> gpurun_out/bench_n8_m24.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_n[48]_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "%.4e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"], d["clocks"]["reasons"], d["config"]["setup_s"])
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/bench_n8_m111.err
