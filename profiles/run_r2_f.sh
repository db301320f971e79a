#!/bin/bash
# round 2, GPU call F (1 GPU): full GPU test-suite, smoke, bench N=1 at the driver's flags, launch list + full ncu capture (traffic)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2f; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "rel-L2|passed|failed" $O/pytest_gpu.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err; echo "ref rc=$?"; cut -c1-250 $O/bench_ref_n1.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; cut -c1-250 $O/bench_n1.json
timeout 600 ncu --kernel-name regex:saa_k_step --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_m111.csv python bench.py --no-cpu-baseline --no-also --steps 20 --warmup 5 --spin-ms 5 --repeats 5 --e2e-steps 2 > $O/ncu_m111.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --kernel-name regex:saa_k_step --set full --clock-control none --import-source on --launch-skip 30 -c 1 -o $O/ncu_full_saa_k_step_m111 python bench.py --no-cpu-baseline --no-also --steps 20 --warmup 5 --spin-ms 5 --repeats 5 --e2e-steps 2 > $O/ncu_full_m111.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --kernel-name regex:saa_k_step --set full --clock-control none --import-source on --launch-skip 30 -c 1 -o $O/ncu_full_saa_k_step_m24 python bench.py --refine 24 --no-cpu-baseline --no-also --steps 20 --warmup 5 --spin-ms 5 --repeats 5 --e2e-steps 2 > $O/ncu_full_m24.log 2>&1; echo "ncu full m24 rc=$?"
