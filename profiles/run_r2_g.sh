#!/bin/bash
# round 2, GPU call G (1 GPU): full GPU tests, schedule variants of the step kernel after the mass-stream change, K5 occupancy variants,
# reference-exact host-assembled matrix at m = 24
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2g; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "rel-L2|passed|failed" $O/pytest_gpu.log | tail -8
Q="--no-cpu-baseline --no-also --e2e-steps 3 --spin-ms 200"
for v in 0 1 2 3 4 5; do SAA_KVARIANT=$v timeout 300 python bench.py --refine 24 $Q > $O/var_m24_$v.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/var_m24_$v.json')); print('m24 variant $v', d['value'], d['ms_per_step'], d['roofline']['frac'])"; done
for v in 1 2 4 5; do SAA_KVARIANT=$v timeout 300 python bench.py --refine 65 $Q > $O/var_m65_$v.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/var_m65_$v.json')); print('m65 variant $v', d['value'], d['ms_per_step'], d['roofline']['frac'])"; done
for b in 3 4 5 6; do SAA_MF_MINB=$b timeout 300 python bench.py --kernel matfree --refine 24 $Q > $O/mf_m24_minb$b.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/mf_m24_minb$b.json')); print('matfree minb $b', d['value'], d['ms_per_step'], d['matfree']['rel_l2_vs_assembled'])"; done
timeout 900 python bench.py --refine 24 --setup host --no-cpu-baseline --no-also > $O/bench_n1_m24_hostasm.json 2> $O/bench_n1_m24_hostasm.err; python -c "
import json; d=json.load(open('$O/bench_n1_m24_hostasm.json')); print('host-assembled m24', d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['nnz_per_row'])"
