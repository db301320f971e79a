#!/bin/bash
# round 2, GPU call J (2 GPUs): single release fence for all arrival flags — parity + exchange latency
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2j; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_process_per_gpu or two_processes_sharing" > $O/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_sel.log
PORT=29561
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
show() { python -c "
import json; d=json.load(open('$1')); print('$1', d['value'], d['ms_per_step'], d['config']['ms_per_step_without_exchange'], {k:v.get('bit_identical') for k,v in d.get('parity',{}).items()})"; }
timeout 600 $TR $PORT bench.py --gpus 2 --refine 12 --steps 2000 --sync-avoid off --no-also > $O/bench_n2_m12.json 2> $O/bench_n2_m12.err; show $O/bench_n2_m12.json
timeout 600 $TR $((PORT+1)) bench.py --gpus 2 --refine 24 --steps 2000 --sync-avoid off --no-also > $O/bench_n2_m24.json 2> $O/bench_n2_m24.err; show $O/bench_n2_m24.json
timeout 600 $TR $((PORT+2)) bench.py --gpus 2 --steps 20 --warmup 5 --partition blocks --blocks 1x2x1 --sync-avoid off --no-also > $O/bench_n2_ycut.json 2> $O/bench_n2_ycut.err; show $O/bench_n2_ycut.json
timeout 600 $TR $((PORT+3)) bench.py --gpus 2 --steps 20 --warmup 5 --sync-avoid off --no-also > $O/bench_n2_m111.json 2> $O/bench_n2_m111.err; show $O/bench_n2_m111.json
