#!/bin/bash
# final pass A (1 GPU): full gpu test-suite, default bench, 104M-DOF bench, reference arm, ncu evidence of the step kernel
set -x
mkdir -p gpurun_out/finalA
O=gpurun_out/finalA
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -2 $O/bench_n1.err
python bench.py --refine 111 --no-cpu-baseline > $O/bench_n1_m111.json 2> $O/bench_n1_m111.err
python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py --setup host --no-cpu-baseline --no-also > $O/bench_n1_hostasm.json 2> $O/bench_n1_hostasm.err
CMD="python bench.py --launch per_step --steps 20 --warmup 5 --e2e-steps 3 --no-cpu-baseline --no-also"
$CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
$CMD > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:saa_k_step -s 10 -c 3 -o $O/prof_step $CMD > $O/ncu_full.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/finalA/bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.4e"%d["value"], "ms/step %.5f"%d["ms_per_step"], "frac", d.get("roofline",{}).get("frac"), "e2e %.3e"%d["e2e"]["value"], (d.get("also") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "FAILED", e)
PY
