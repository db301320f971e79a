#!/bin/bash
# round 2, GPU call M (1 GPU): pipelined host call with copy-only copy streams; the driver's N = 1 command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2m; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_processes_sharing or page_locked or pipelined or step_host" > $O/pytest_host.log 2>&1; echo "pytest host rc=$?"; tail -4 $O/pytest_host.log
show() { python -c "
import json; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], json.dumps(d['e2e'])[:420])"; }
S=$(date +%s); timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1_s20.json 2> $O/bench_n1_s20.err; echo "bench rc=$? wall=$(( $(date +%s) - S )) s"; show $O/bench_n1_s20.json
timeout 600 python bench.py --refine 24 --no-cpu-baseline --no-also --steps 20 --warmup 5 > $O/bench_n1_m24.json 2> $O/bench_n1_m24.err; show $O/bench_n1_m24.json
timeout 600 python bench.py --refine 65 --no-cpu-baseline --no-also --steps 20 --warmup 5 > $O/bench_n1_m65.json 2> $O/bench_n1_m65.err; show $O/bench_n1_m65.json
S=$(date +%s); timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err; echo "ref rc=$? wall=$(( $(date +%s) - S )) s"; cut -c1-600 $O/bench_ref_n1.json
