#!/bin/bash
# third GPU pass: device set-up tests, bench N=1 on the device path (+ 21M DOF), 104M-DOF single-GPU probe
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_device_setup.py -x -q > gpurun_out/pytest_devsetup.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_devsetup.log
tail -25 gpurun_out/pytest_devsetup.log
python bench.py > gpurun_out/bench_n1_dev.json 2> gpurun_out/bench_n1_dev.err; tail -c 3500 gpurun_out/bench_n1_dev.json; tail -5 gpurun_out/bench_n1_dev.err
python bench.py --m 111 --steps 300 --warmup 10 --e2e-steps 5 --no-cpu-baseline > gpurun_out/bench_n1_m111.json 2> gpurun_out/bench_n1_m111.err; tail -c 2500 gpurun_out/bench_n1_m111.json; tail -5 gpurun_out/bench_n1_m111.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
