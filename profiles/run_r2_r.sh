#!/bin/bash
# round 2, GPU call R (1 GPU): full pipeline test on genuine HDF5 result files
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2r; mkdir -p $O
timeout 75 python -m pytest -x -q tests/test_lstm.py::test_pipeline_with_surrogates_trained_by_the_unmodified_reference_script > $O/pytest_pipeline.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_pipeline.log
