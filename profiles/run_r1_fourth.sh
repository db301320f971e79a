#!/bin/bash
# fourth GPU pass (2 GPUs): device set-up tests; N=2 strong scaling on the 21M-DOF mesh, peer vs NCCL
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_device_setup.py -x -q > gpurun_out/pytest_devsetup.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_devsetup.log
tail -25 gpurun_out/pytest_devsetup.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --transport peer > gpurun_out/bench_n2_m65_peer.json 2> gpurun_out/bench_n2_m65_peer.err; tail -c 2200 gpurun_out/bench_n2_m65_peer.json; tail -5 gpurun_out/bench_n2_m65_peer.err
$TR bench.py --gpus 2 --transport nccl > gpurun_out/bench_n2_m65_nccl.json 2> gpurun_out/bench_n2_m65_nccl.err; tail -c 2200 gpurun_out/bench_n2_m65_nccl.json; tail -5 gpurun_out/bench_n2_m65_nccl.err
$TR bench.py --gpus 2 --m 24 --transport peer --no-cpu-baseline > gpurun_out/bench_n2_m24_peer.json 2> gpurun_out/bench_n2_m24_peer.err; tail -c 600 gpurun_out/bench_n2_m24_peer.json
