#!/bin/bash
# round 2, GPU call H (2 GPUs): persistent synchronised loop — parity (dist worker) + small-shard timing vs the graph-replayed fused step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2h; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_process_per_gpu and peer" > $O/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_sel.log
PORT=29541
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
i=0
for L in graph persistent; do for M in 24 12; do i=$((i+1)); timeout 600 $TR $((PORT+i)) bench.py --gpus 2 --refine $M --launch $L --steps 2000 --sync-avoid off --no-also > $O/bench_n2_m${M}_$L.json 2> $O/bench_n2_m${M}_$L.err; python -c "
import json; d=json.load(open('$O/bench_n2_m${M}_$L.json')); print('m$M $L', d['value'], d['ms_per_step'], d['config']['ms_per_step_without_exchange'], d['gpu_launches'])"; done; done
