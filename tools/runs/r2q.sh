#!/bin/bash
# round-2 GPU run Q (1 GPU): Horizon patch stage after the compact-array / ballot changes: suite (bit-exact extractor), stamps, timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2q_tests.log
LILIOM_DEBUG_TIMING=1 timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-dense-probe --e2e sequential > gpurun_out/r2q_dbg.json 2> gpurun_out/r2q_dbg.err
timeout 400 python tools/ab_variants.py 200 3 3 > gpurun_out/r2q_ab.log 2>&1
tail -3 gpurun_out/r2q_tests.log; grep "coop" gpurun_out/r2q_dbg.err | tail -2; cat gpurun_out/r2q_ab.log
