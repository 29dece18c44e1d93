#!/bin/bash
# round-2 GPU run D (1 GPU): suite + dense sweep + bench after the dense-kernel v2 / VoxelGrid phase-2 / GN-tail changes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log
timeout 400 python tools/knn_sweep.py > gpurun_out/r2d_sweep.log 2>&1
timeout 200 python tools/knn_once.py 10000000 hdl > gpurun_out/r2d_once_hdl10m.log 2>&1
LILIOM_DEBUG_TIMING=1 timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-dense-probe --e2e sequential > gpurun_out/r2d_dbg.json 2> gpurun_out/r2d_dbg.err
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -4 gpurun_out/r2d_tests.log; cat gpurun_out/r2d_sweep.log gpurun_out/r2d_once_hdl10m.log; grep "coop\|persistent" gpurun_out/r2d_dbg.err | tail -4; cat gpurun_out/r2d_bench.json | cut -c1-400
