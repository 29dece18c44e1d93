#!/bin/bash
# round-2 GPU run N (1 GPU): node mirror on liliom_map_update, streamed workload with the incremental map leg, real-size lifecycle
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2n_tests.log
timeout 600 python bench.py --workload stream --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_stream1.json 2> gpurun_out/r2n_stream1.err
for mode in 0 1; do echo "mode=$mode: $(timeout 300 python tools/stream_bench.py 60 $mode 2>&1 | tail -1)" >> gpurun_out/r2n_stream_real.log; done
tail -3 gpurun_out/r2n_tests.log; cat gpurun_out/r2n_stream_real.log; python -c "
import json; j=json.loads([l for l in open('gpurun_out/r2n_stream1.json') if l.startswith('{')][-1]); print(j['value'], j['ms_per_step'], j.get('step_breakdown_ms'), j['pose_err_m']); print(j.get('incremental_map'))"; tail -3 gpurun_out/r2n_stream1.err
