#!/bin/bash
# round-2 GPU run K (1 GPU): suite after the rebuild changes; real-size streamed lifecycle with and without the cooperative
# map filter; streamed 10 M workload phases
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2k_tests.log
for m in 0 1; do
  for mode in 0 1; do
    echo "MAP_COOP=$m mode=$mode: $(LILIOM_MAP_COOP=$m timeout 300 python tools/stream_bench.py 60 $mode 2>&1 | tail -1)" >> gpurun_out/r2k_stream_real.log
  done
done
LILIOM_DEBUG_TIMING=1 timeout 600 python bench.py --workload stream --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_dbg1.json 2> gpurun_out/r2k_dbg1.err
tail -3 gpurun_out/r2k_tests.log; cat gpurun_out/r2k_stream_real.log; grep "map_rebuild" gpurun_out/r2k_dbg1.err | tail -2; python -c "
import json; j=json.loads([l for l in open('gpurun_out/r2k_dbg1.json') if l.startswith('{')][-1]); print(j['value'], j['ms_per_step'], j.get('step_breakdown_ms'), j['pose_err_m'])"
