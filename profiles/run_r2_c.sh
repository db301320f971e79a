#!/bin/bash
# round 2, GPU call C (8 GPUs): P = 4 / 8 parity tests (golden + mid-size METIS), bench N = 8 (slabs + sync-avoiding; 2x2x2 blocks), N = 4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_process_per_gpu and (P4 or P8 or np4 or np8)" > $O/pytest_peer8.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_peer8.log
PORT=29521
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port $PORT bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"; cut -c1-300 $O/bench_n8.json; tail -3 $O/bench_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port $((PORT+1)) bench.py --gpus 8 --steps 20 --warmup 5 --partition blocks --sync-avoid off --no-also > $O/bench_n8_blocks.json 2> $O/bench_n8_blocks.err; echo "bench n8 blocks rc=$?"; cut -c1-300 $O/bench_n8_blocks.json; tail -3 $O/bench_n8_blocks.err
timeout 600 $TR --nproc-per-node 4 --master-port $((PORT+2)) bench.py --gpus 4 --steps 20 --warmup 5 --sync-avoid off > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench n4 rc=$?"; cut -c1-300 $O/bench_n4.json; tail -3 $O/bench_n4.err
timeout 400 $TR --nproc-per-node 8 --master-port $((PORT+3)) bench.py --gpus 8 --refine 24 --partition metis --sync-avoid off --no-also > $O/bench_n8_m24_metis.json 2> $O/bench_n8_m24_metis.err; echo "bench n8 m24 metis rc=$?"; cut -c1-300 $O/bench_n8_m24_metis.json; tail -3 $O/bench_n8_m24_metis.err
