#!/bin/bash
# round-2 GPU run A: full GPU suite, bench line, dense sweep
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> gpurun_out/r2a_bench.err
timeout 400 python tools/knn_sweep.py > gpurun_out/r2a_sweep.log 2>&1; echo "sweep rc=$?" >> gpurun_out/r2a_sweep.log
tail -5 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_bench.json; tail -12 gpurun_out/r2a_sweep.log
