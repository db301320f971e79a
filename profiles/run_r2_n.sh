#!/bin/bash
# round 2, GPU call N (2 GPUs): PCIe duplex ceiling; pipelined synchronised host call over the peer and NCCL transports
# (one process per GPU, golden + mid-size fixtures); the driver's N = 2 command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2n; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 120 python profiles/pcie_probe.py 836 > $O/pcie_probe.json 2> $O/pcie_probe.err; cat $O/pcie_probe.json
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_process_per_gpu" > $O/pytest_2gpu.log 2>&1; echo "pytest 2gpu rc=$?"; tail -4 $O/pytest_2gpu.log
show() { python -c "
import json; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], json.dumps(d['e2e'])[:420], d.get('parity'))"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
S=$(date +%s); timeout 900 $TR 29571 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_s20.json 2> $O/bench_n2_s20.err; echo "bench n2 rc=$? wall=$(( $(date +%s) - S )) s"; show $O/bench_n2_s20.json; tail -3 $O/bench_n2_s20.err
timeout 600 $TR 29572 bench.py --gpus 2 --steps 20 --warmup 5 --partition blocks --blocks 1x2x1 --sync-avoid off --no-also > $O/bench_n2_ycut.json 2> $O/bench_n2_ycut.err; show $O/bench_n2_ycut.json
