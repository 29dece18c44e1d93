#!/bin/bash
# round-2 sanitizer pass (1 GPU, the round's last GPU seconds): compute-sanitizer memcheck over the tests of this round's new kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export LILIOM_ASSUME_GPU=1
timeout 48 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -k "walk_paths or box_escape or lifecycle or rot_extract_bit_exact" -x -q > gpurun_out/r2san_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2san_memcheck.log
tail -8 gpurun_out/r2san_memcheck.log
