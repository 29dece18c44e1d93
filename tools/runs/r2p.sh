#!/bin/bash
# round-2 GPU run P (1 GPU): loop-closure ICP (f4) test, whole suite, Horizon patch-stage stamps
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_widen.py -m gpu -x -q -k "icp" > gpurun_out/r2p_icp.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_icp.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2p_tests.log
LILIOM_DEBUG_TIMING=1 timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-dense-probe --e2e sequential > gpurun_out/r2p_dbg.json 2> gpurun_out/r2p_dbg.err
tail -25 gpurun_out/r2p_icp.log | cut -c1-300; tail -3 gpurun_out/r2p_tests.log; grep "hz_coop" gpurun_out/r2p_dbg.err | tail -2
