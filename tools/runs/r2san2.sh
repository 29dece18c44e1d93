#!/bin/bash
# round-2 sanitizer pass 2 (1 GPU): memcheck over the parity / variants / widen tiers, racecheck on the ring-walk test
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export LILIOM_ASSUME_GPU=1
timeout 55 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py tests/test_gpu_widen.py -x -q > gpurun_out/r2san2_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2san2_memcheck.log
tail -5 gpurun_out/r2san2_memcheck.log
timeout 30 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -k "walk_paths" -x -q > gpurun_out/r2san2_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2san2_racecheck.log
tail -6 gpurun_out/r2san2_racecheck.log
