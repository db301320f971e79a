#!/bin/bash
# round 2, GPU call A (1 GPU): tests, mid-size fixtures, first bench with the new timing protocol, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2a; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_fused_peer_step_two_processes_sharing_this_gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 300 python oracle/gen_golden_mid.py $O/golden > $O/gen_mid.log 2>&1; echo "gen rc=$?"; tail -4 $O/gen_mid.log
timeout 420 python -m pytest tests/test_gpu_parity.py -q -k two_processes_sharing > $O/pytest_shared_gpu.log 2>&1; echo "shared-gpu rc=$?"; tail -15 $O/pytest_shared_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1_s20.json 2> $O/bench_n1_s20.err; echo "bench rc=$?"; cat $O/bench_n1_s20.json | cut -c1-1500
timeout 600 python bench.py --no-cpu-baseline --no-also > $O/bench_n1_default.json 2> $O/bench_n1_default.err; echo "bench default rc=$?"; cut -c1-600 $O/bench_n1_default.json
timeout 600 python bench.py --refine 24 --no-cpu-baseline --no-also --steps 20 --warmup 5 > $O/bench_n1_m24_s20.json 2> $O/bench_n1_m24.err; cut -c1-400 $O/bench_n1_m24_s20.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_m24.csv python bench.py --refine 24 --no-cpu-baseline --no-also --steps 20 --warmup 5 --spin-ms 5 --repeats 5 --e2e-steps 3 > $O/ncu_m24.log 2>&1; echo "ncu rc=$?"
