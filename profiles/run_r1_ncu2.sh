#!/bin/bash
# ncu evidence for the node-block step kernel (1 GPU): launch list + full capture
set -x
mkdir -p gpurun_out
CMD="python bench.py --launch per_step --steps 20 --warmup 5 --e2e-steps 3 --no-cpu-baseline --no-also"
$CMD > gpurun_out/plain_nb.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_nb.csv $CMD > gpurun_out/ncu_list_nb.log 2>&1
$CMD > gpurun_out/plain_nb2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:saa_k_step -s 10 -c 3 -o gpurun_out/prof_step_nb $CMD > gpurun_out/ncu_full_nb.log 2>&1
tail -3 gpurun_out/ncu_full_nb.log
