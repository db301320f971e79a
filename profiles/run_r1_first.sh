#!/bin/bash
# first GPU pass of round 1: parity tests, smoke, bench, ncu launch list + full capture of the step kernel
set -x
mkdir -p gpurun_out
export SAA_BENCH_CACHE=/tmp/saa_cache
nvidia-smi -L
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; tail -c 3000 gpurun_out/bench_graph.json
python bench.py --launch persistent --no-cpu-baseline > gpurun_out/bench_persistent.json 2> gpurun_out/bench_persistent.err; tail -c 1500 gpurun_out/bench_persistent.json
python bench.py --launch per_step --no-cpu-baseline > gpurun_out/bench_perstep.json 2> gpurun_out/bench_perstep.err; tail -c 1500 gpurun_out/bench_perstep.json
CMD="python bench.py --launch per_step --steps 20 --warmup 5 --e2e-steps 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:saa_k_step -s 10 -c 3 -o gpurun_out/prof_step $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
