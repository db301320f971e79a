#!/bin/bash
# final pass B (8 GPUs): P=4 / P=8 transport parity; strong scaling N=8 and N=4 on 21M and 104M DOF; sync-avoiding; balance
set -x
mkdir -p gpurun_out/finalB
O=gpurun_out/finalB
python -m pytest tests/test_gpu_parity.py -q -k "one_process_per_gpu" > $O/pytest_peer8.log 2>&1; echo "pytest exit $?" >> $O/pytest_peer8.log; tail -3 $O/pytest_peer8.log
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
$T8 bench.py --gpus 8 --steps 4000 --sync-avoid 0,10,50 > $O/bench_n8_m65.json 2> $O/bench_n8_m65.err
$T8 bench.py --gpus 8 --refine 111 --steps 2000 > $O/bench_n8_m111.json 2> $O/bench_n8_m111.err
$T8 bench.py --gpus 8 --steps 4000 --balance > $O/bench_n8_m65_bal.json 2> $O/bench_n8_m65_bal.err
$T8 bench.py --gpus 8 --steps 4000 --transport nccl > $O/bench_n8_m65_nccl.json 2> $O/bench_n8_m65_nccl.err
$T4 bench.py --gpus 4 --steps 3000 > $O/bench_n4_m65.json 2> $O/bench_n4_m65.err
$T4 bench.py --gpus 4 --refine 111 --steps 1500 > $O/bench_n4_m111.json 2> $O/bench_n4_m111.err
$T8 bench.py --gpus 8 --refine 24 > $O/bench_n8_m24.json 2> $O/bench_n8_m24.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/finalB/bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.4e"%d["value"], "ms/step %.5f"%d["ms_per_step"], "noexch %.5f"%d["config"]["ms_per_step_without_exchange"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"], d["config"].get("balance"))
        if "sync_avoiding" in d: print("   ", [(r["resync_every"], "%.4e"%r["value"]) for r in d["sync_avoiding"]["runs"]])
    except Exception as e: print(f, "FAILED", e)
PY
