#!/bin/bash
# eighth GPU pass (2 GPUs): sync-avoiding driver test, N=2 bench with the node-block layout, sync-avoiding timing, balancing
set -x
mkdir -p gpurun_out
python -m pytest tests/test_lstm.py tests/test_gpu_parity.py -x -q -m gpu -k "online_predictor or one_process_per_gpu or sync_avoiding" > gpurun_out/pytest_e.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_e.log
tail -6 gpurun_out/pytest_e.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --no-cpu-baseline --sync-avoid 0,10,50 > gpurun_out/bench_n2_nb.json 2> gpurun_out/bench_n2_nb.err; tail -4 gpurun_out/bench_n2_nb.err
$TR bench.py --gpus 2 --no-cpu-baseline --balance > gpurun_out/bench_n2_nb_bal.json 2> gpurun_out/bench_n2_nb_bal.err; tail -4 gpurun_out/bench_n2_nb_bal.err
python - <<'PY'
import json
for f in ["bench_n2_nb","bench_n2_nb_bal"]:
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f, "value %.4e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"], d["config"].get("balance"))
        if "sync_avoiding" in d: print(json.dumps(d["sync_avoiding"]))
    except Exception as e: print(f, "FAILED", e)
PY
