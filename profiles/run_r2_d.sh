#!/bin/bash
# round 2, GPU call D (2 GPUs): overlapped shared-row units — parity tests + exchange cost at m = 65 / 111
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2e; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/pytest_sel.log 2>&1; echo "pytest rc=$?"; grep -E "rel-L2|passed|failed" $O/pytest_sel.log | tail -8
PORT=29531
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 600 $TR $PORT bench.py --gpus 2 --refine 65 --sync-avoid off --no-also > $O/bench_n2_m65.json 2> $O/bench_n2_m65.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$O/bench_n2_m65.json')); print(d['value'], d['ms_per_step'], d['config']['ms_per_step_without_exchange'], d['parity']['cross_path'])"
timeout 600 $TR $((PORT+1)) bench.py --gpus 2 --steps 20 --warmup 5 --sync-avoid off --no-also > $O/bench_n2_m111.json 2> $O/bench_n2_m111.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$O/bench_n2_m111.json')); print(d['value'], d['ms_per_step'], d['config']['ms_per_step_without_exchange'], d['parity']['cross_path'])"
timeout 600 $TR $((PORT+2)) bench.py --gpus 2 --steps 20 --warmup 5 --partition blocks --blocks 1x2x1 --sync-avoid off --no-also > $O/bench_n2_ycut.json 2> $O/bench_n2_ycut.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$O/bench_n2_ycut.json')); print(d['value'], d['ms_per_step'], d['config']['ms_per_step_without_exchange'], d['parity']['cross_path'])"
timeout 600 $TR $((PORT+3)) bench.py --gpus 2 --refine 24 --sync-avoid off --no-also > $O/bench_n2_m24.json 2> $O/bench_n2_m24.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$O/bench_n2_m24.json')); print(d['value'], d['ms_per_step'], d['config']['ms_per_step_without_exchange'], d['parity']['cross_path'])"
