#!/bin/bash
# round-2 GPU run M (1 GPU): incremental map (f2) parity tests + whole suite
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "incremental" > gpurun_out/r2m_inc.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_inc.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_tests.log
tail -30 gpurun_out/r2m_inc.log | cut -c1-300; tail -3 gpurun_out/r2m_tests.log
