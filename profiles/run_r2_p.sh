#!/bin/bash
# round 2, GPU call P (1 GPU): final state — full GPU suite (shim with the peer transport between two processes sharing
# the GPU), smoke, the driver's N = 1 command (e2e of the also configs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2p; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -6 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
S=$(date +%s); timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1_s20.json 2> $O/bench_n1_s20.err; echo "bench rc=$? wall=$(( $(date +%s) - S )) s"
python -c "
import json; d=json.loads(open('$O/bench_n1_s20.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], json.dumps(d['e2e'])[:300]); print([(a.get('value'), a.get('roofline_frac'), a.get('e2e',{}).get('value'), a.get('e2e',{}).get('bit_identical_to_resident_steps'), a.get('error')) for a in d['also']])"
tail -3 $O/bench_n1_s20.err
