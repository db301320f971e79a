#!/bin/bash
# round 2, GPU call Q (1 GPU, last seconds of the budget): the result files of the drivers are genuine HDF5 now
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2q; mkdir -p $O
timeout 110 python -m pytest -x -q "tests/test_tools_shim.py::test_data_prepare_shaped_driver_matches_reference_history" tests/test_lstm.py::test_online_predictor_shaped_driver_two_processes > $O/pytest_results.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_results.log
