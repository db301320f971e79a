#!/bin/bash
# round 2, GPU call K (1 GPU): pipelined host call (upload / step / download overlapped), pinned result pool — parity + e2e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2k; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or page_locked or step_host" > $O/pytest_host.log 2>&1; echo "pytest host rc=$?"; tail -25 $O/pytest_host.log
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_processes_sharing" > $O/pytest_shared.log 2>&1; echo "pytest shared rc=$?"; tail -25 $O/pytest_shared.log
show() { python -c "
import json; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], json.dumps(d['e2e']))"; }
timeout 600 python bench.py --no-cpu-baseline --no-also --steps 20 --warmup 5 > $O/bench_n1_m111.json 2> $O/bench_n1_m111.err; echo "bench rc=$?"; show $O/bench_n1_m111.json; tail -5 $O/bench_n1_m111.err
SAA_STEP_HOST_PIPELINE=0 timeout 600 python bench.py --no-cpu-baseline --no-also --steps 20 --warmup 5 > $O/bench_n1_m111_nopipe.json 2> $O/bench_n1_m111_nopipe.err; show $O/bench_n1_m111_nopipe.json
timeout 600 python bench.py --refine 24 --no-cpu-baseline --no-also --steps 20 --warmup 5 > $O/bench_n1_m24.json 2> $O/bench_n1_m24.err; show $O/bench_n1_m24.json
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_fused_peer_step_two_processes_sharing_this_gpu --deselect tests/test_gpu_parity.py::test_mid_size_partition_two_processes_sharing_this_gpu > $O/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -5 $O/pytest_gpu.log
