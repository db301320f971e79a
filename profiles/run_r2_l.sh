#!/bin/bash
# round 2, GPU call L (1 GPU): pipelined host call with <= 32 chunks, synchronised host call on the mid-size partition
# (two processes sharing the GPU), full GPU suite, the driver's N = 1 command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2l; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_processes_sharing or page_locked" > $O/pytest_shared.log 2>&1; echo "pytest shared rc=$?"; tail -25 $O/pytest_shared.log
show() { python -c "
import json; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], json.dumps(d['e2e'])[:420])"; }
/usr/bin/time -v timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1_s20.json 2> $O/bench_n1_s20.err; echo "bench rc=$?"; show $O/bench_n1_s20.json; grep -E "Elapsed|Maximum resident" $O/bench_n1_s20.err
timeout 600 python bench.py --refine 24 --no-cpu-baseline --no-also --steps 20 --warmup 5 > $O/bench_n1_m24.json 2> $O/bench_n1_m24.err; show $O/bench_n1_m24.json
timeout 600 python bench.py --refine 65 --no-cpu-baseline --no-also --steps 20 --warmup 5 > $O/bench_n1_m65.json 2> $O/bench_n1_m65.err; show $O/bench_n1_m65.json
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -5 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
