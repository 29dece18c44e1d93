#!/bin/bash
# round-2 GPU run O (1 GPU): scalar eigen-solver + ballot list in the Horizon patch stage; merge search bounded per warp
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2o_tests.log
LILIOM_DEBUG_TIMING=1 timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-dense-probe --e2e sequential > gpurun_out/r2o_dbg.json 2> gpurun_out/r2o_dbg.err
timeout 400 python tools/ab_variants.py 150 3 3 > gpurun_out/r2o_ab.log 2>&1
timeout 600 python bench.py --workload stream --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_stream1.json 2> gpurun_out/r2o_stream1.err
tail -3 gpurun_out/r2o_tests.log; grep "coop" gpurun_out/r2o_dbg.err | tail -2; cat gpurun_out/r2o_ab.log; python -c "
import json; j=json.loads([l for l in open('gpurun_out/r2o_stream1.json') if l.startswith('{')][-1]); print(j['value'], j['ms_per_step'], j.get('step_breakdown_ms')); print(j.get('incremental_map'))"
