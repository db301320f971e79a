#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_nb.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_nb.log
tail -8 gpurun_out/pytest_gpu_nb.log
python bench.py > gpurun_out/bench_n1_nb.json 2> gpurun_out/bench_n1_nb.err; tail -c 3800 gpurun_out/bench_n1_nb.json; tail -3 gpurun_out/bench_n1_nb.err
python bench.py --setup host --no-cpu-baseline --no-also > gpurun_out/bench_n1_nb_host.json 2> gpurun_out/bench_n1_nb_host.err; tail -c 1500 gpurun_out/bench_n1_nb_host.json
