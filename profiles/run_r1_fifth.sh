#!/bin/bash
# fifth GPU pass (2 GPUs): multi-GPU parity with the fused single-launch step; A/B fused vs three kernels
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "one_process_per_gpu" > gpurun_out/pytest_peer.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_peer.log
tail -5 gpurun_out/pytest_peer.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_n2_m65_fused.json 2> gpurun_out/bench_n2_m65_fused.err; tail -c 700 gpurun_out/bench_n2_m65_fused.json | head -c 300; tail -3 gpurun_out/bench_n2_m65_fused.err
SAA_PEER_FUSED=0 $TR bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_n2_m65_3k.json 2> gpurun_out/bench_n2_m65_3k.err
$TR bench.py --gpus 2 --refine 24 > gpurun_out/bench_n2_m24_fused.json 2> gpurun_out/bench_n2_m24_fused.err
SAA_PEER_FUSED=0 $TR bench.py --gpus 2 --refine 24 > gpurun_out/bench_n2_m24_3k.json 2> gpurun_out/bench_n2_m24_3k.err
python - <<'PY'
import json
for f in ["bench_n2_m65_fused","bench_n2_m65_3k","bench_n2_m24_fused","bench_n2_m24_3k"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, "%.4e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "launches", d["gpu_launches"], "e2e %.3e"%d["e2e"]["value"])
    except Exception as e:
        print(f, "FAILED", e)
PY
