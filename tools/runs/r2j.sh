#!/bin/bash
# round-2 GPU run J (2 GPUs): phase clocks of the sharded map maintenance, then the GPU suite (new undistortion test) on GPU 0
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
LILIOM_DEBUG_TIMING=1 timeout 600 $TR bench.py --gpus 2 --steps 4 --warmup 3 --no-extra-legs > gpurun_out/r2j_dbg2.json 2> gpurun_out/r2j_dbg2.err
LILIOM_DEBUG_TIMING=1 timeout 600 python bench.py --workload stream --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_dbg1.json 2> gpurun_out/r2j_dbg1.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2j_tests.log
bash tools/runs/r2i.sh > /dev/null 2>&1
grep "map_rebuild" gpurun_out/r2j_dbg2.err | tail -4; grep "map_rebuild" gpurun_out/r2j_dbg1.err | tail -3; tail -3 gpurun_out/r2j_tests.log; cat gpurun_out/r2i_lanes.log
