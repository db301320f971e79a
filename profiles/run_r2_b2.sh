#!/bin/bash
# round 2, GPU call B2 (2 GPUs): 2-GPU parity tests (golden + mid-size), bench N=2 (slabs, sync-avoiding leg; y-cut blocks)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2b2; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_matfree.py tests/test_gpu_parity.py -m gpu -x -q -s -k "matfree or mid_fixture or one_process_per_gpu" > $O/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_sel.log
PORT=29511
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 900 $TR $PORT bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; cut -c1-400 $O/bench_n2.json; tail -5 $O/bench_n2.err
timeout 600 $TR $((PORT+1)) bench.py --gpus 2 --steps 20 --warmup 5 --partition blocks --blocks 1x2x1 --sync-avoid off --no-also > $O/bench_n2_ycut.json 2> $O/bench_n2_ycut.err; echo "bench n2 ycut rc=$?"; cut -c1-400 $O/bench_n2_ycut.json; tail -3 $O/bench_n2_ycut.err
timeout 600 $TR $((PORT+2)) bench.py --gpus 2 --refine 24 --partition metis --sync-avoid off --no-also > $O/bench_n2_m24_metis.json 2> $O/bench_n2_m24_metis.err; echo "bench n2 metis rc=$?"; cut -c1-300 $O/bench_n2_m24_metis.json; tail -3 $O/bench_n2_m24_metis.err
