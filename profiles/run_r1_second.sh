#!/bin/bash
# second GPU pass (2 GPUs): full gpu test-suite incl. Tools shim and one-process-per-GPU transports; N=2 bench
set -x
mkdir -p gpurun_out
export SAA_BENCH_CACHE=/tmp/saa_cache
nvidia-smi -L; nvidia-smi topo -m | head -8
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu2.log
tail -15 gpurun_out/pytest_gpu2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --transport peer > gpurun_out/bench_n2_peer.json 2> gpurun_out/bench_n2_peer.err; tail -c 1800 gpurun_out/bench_n2_peer.json; tail -5 gpurun_out/bench_n2_peer.err
$TR bench.py --gpus 2 --transport nccl > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; tail -c 1800 gpurun_out/bench_n2_nccl.json; tail -5 gpurun_out/bench_n2_nccl.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
