#!/bin/bash
# round-2 GPU run W (1 GPU): per-ring stage cycles of k_rot_ring; 8-lane vs 16-lane x 2 rounds on the 3.1k-query scans
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
LILIOM_DEBUG_TIMING=1 timeout 300 python bench.py --workload rot --steps 3 --warmup 2 --no-cpu-baseline --no-dense-probe --e2e sequential > gpurun_out/r2w_dbg_rot.json 2> gpurun_out/r2w_dbg_rot.err; grep "k_rot_ring" gpurun_out/r2w_dbg_rot.err | tail -4
for lanes in 0 8 0 8; do
  if [ $lanes = 0 ]; then unset LILIOM_KNN_LANES; else export LILIOM_KNN_LANES=$lanes; fi
  timeout 300 python bench.py --workload rot --steps 100 --warmup 5 --no-cpu-baseline --no-dense-probe --e2e sequential > gpurun_out/r2w_rot_l$lanes.json 2> gpurun_out/r2w_rot_l$lanes.err
  python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/r2w_rot_l$lanes.json') if l.startswith('{')][-1])
print('lanes=$lanes', 'value',round(j['value'],1),'ms',round(j['ms_per_step'],4),'GN us/pass',round(j['roofline']['us_per_launch'],2),'queries',j['run']['queries_per_scan'],'pose_err',j['pose_err_m'])
PY
done
unset LILIOM_KNN_LANES
