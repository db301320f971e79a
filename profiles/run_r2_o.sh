#!/bin/bash
# round 2, GPU call O (4 GPUs): pipelined synchronised host call with two neighbours per rank / METIS partitions with
# nodes held by >= 3 ranks (peer and NCCL transports), the driver's N = 4 command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2o; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_process_per_gpu and (P4 or np4)" > $O/pytest_4gpu.log 2>&1; echo "pytest 4gpu rc=$?"; tail -4 $O/pytest_4gpu.log
show() { python -c "
import json; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], json.dumps(d['e2e'])[:420], {k:v.get('bit_identical') for k,v in d.get('parity',{}).items()})"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port"
S=$(date +%s); timeout 600 $TR 29581 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_n4_s20.json 2> $O/bench_n4_s20.err; echo "bench n4 rc=$? wall=$(( $(date +%s) - S )) s"; show $O/bench_n4_s20.json; tail -3 $O/bench_n4_s20.err
