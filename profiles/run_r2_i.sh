#!/bin/bash
# round 2, GPU call I (8 GPUs): final build — P = 4 / 8 parity tests, strong-scaling points N = 8 / 4 on the 104 M-DOF mesh,
# 2x2x2 blocks, small-shard case (1.13 M DOF on 8 GPUs) graph vs persistent
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2i; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_process_per_gpu and (P4 or P8 or np4 or np8)" > $O/pytest_peer8.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_peer8.log
PORT=29551
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
show() { python -c "
import json; d=json.load(open('$1')); print('$1', d['value'], d['ms_per_step'], d['config']['ms_per_step_without_exchange'], {k:v.get('bit_identical') for k,v in d.get('parity',{}).items()})"; }
timeout 900 $TR --nproc-per-node 8 --master-port $PORT bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "rc=$?"; show $O/bench_n8.json
timeout 600 $TR --nproc-per-node 4 --master-port $((PORT+1)) bench.py --gpus 4 --steps 20 --warmup 5 --sync-avoid off --no-also > $O/bench_n4.json 2> $O/bench_n4.err; echo "rc=$?"; show $O/bench_n4.json
timeout 600 $TR --nproc-per-node 8 --master-port $((PORT+2)) bench.py --gpus 8 --steps 20 --warmup 5 --partition blocks --sync-avoid off --no-also > $O/bench_n8_blocks.json 2> $O/bench_n8_blocks.err; echo "rc=$?"; show $O/bench_n8_blocks.json
timeout 400 $TR --nproc-per-node 8 --master-port $((PORT+3)) bench.py --gpus 8 --refine 24 --steps 2000 --sync-avoid off --no-also > $O/bench_n8_m24_graph.json 2> $O/bench_n8_m24_graph.err; echo "rc=$?"; show $O/bench_n8_m24_graph.json
timeout 400 $TR --nproc-per-node 8 --master-port $((PORT+4)) bench.py --gpus 8 --refine 24 --steps 2000 --launch persistent --sync-avoid off --no-also > $O/bench_n8_m24_persistent.json 2> $O/bench_n8_m24_persistent.err; echo "rc=$?"; show $O/bench_n8_m24_persistent.json
timeout 400 $TR --nproc-per-node 8 --master-port $((PORT+5)) bench.py --gpus 8 --refine 24 --partition metis --sync-avoid off --no-also > $O/bench_n8_m24_metis.json 2> $O/bench_n8_m24_metis.err; echo "rc=$?"; show $O/bench_n8_m24_metis.json
