#!/bin/bash
# round 2, GPU call B1 (1 GPU): mid-size fixtures, K5 tests + bench, launch list of the step kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2b; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 300 python oracle/gen_golden_mid.py $O/golden > $O/gen_mid.log 2>&1; echo "gen rc=$?"; tail -4 $O/gen_mid.log
cp $O/golden/mid_np*.npz tests/golden/
timeout 900 python -m pytest tests/test_gpu_matfree.py tests/test_gpu_parity.py tests/test_lstm.py -m gpu -x -q -s -k "matfree or mid_fixture or step_host_skips or chunked" > $O/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_sel.log
timeout 600 python bench.py --kernel matfree --refine 24 --no-cpu-baseline > $O/bench_matfree_m24.json 2> $O/bench_matfree_m24.err; echo "matfree m24 rc=$?"; cut -c1-300 $O/bench_matfree_m24.json; tail -3 $O/bench_matfree_m24.err
timeout 600 python bench.py --kernel matfree --no-cpu-baseline > $O/bench_matfree_m111.json 2> $O/bench_matfree_m111.err; echo "matfree m111 rc=$?"; cut -c1-300 $O/bench_matfree_m111.json; tail -3 $O/bench_matfree_m111.err
timeout 600 ncu --kernel-name regex:saa_k_step --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_m24.csv python bench.py --refine 24 --no-cpu-baseline --no-also --steps 20 --warmup 5 --spin-ms 5 --repeats 5 --e2e-steps 3 > $O/ncu_m24.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --kernel-name regex:saa_k_step_matfree --set full --clock-control none --import-source on --launch-skip 20 -c 1 -o $O/ncu_full_matfree_m24 python bench.py --kernel matfree --refine 24 --no-cpu-baseline --steps 20 --warmup 5 --spin-ms 5 --repeats 5 --e2e-steps 3 > $O/ncu_full_mf.log 2>&1; echo "ncu full rc=$?"
