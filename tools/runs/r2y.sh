#!/bin/bash
# round-2 GPU run Y (1 GPU): k_rot_ring mask walk — parity tests, stage cycles, rot + stream bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2y_tests.log; tail -5 gpurun_out/r2y_tests.log
LILIOM_DEBUG_TIMING=1 timeout 300 python bench.py --workload rot --steps 3 --warmup 2 --no-cpu-baseline --no-dense-probe --e2e sequential > gpurun_out/r2y_dbg_rot.json 2> gpurun_out/r2y_dbg_rot.err; grep "k_rot_ring" gpurun_out/r2y_dbg_rot.err | tail -2
timeout 300 python bench.py --workload rot --steps 100 --warmup 5 --no-cpu-baseline --no-dense-probe > gpurun_out/r2y_rot.json 2> gpurun_out/r2y_rot.err
timeout 600 python bench.py --workload stream --steps 20 --warmup 3 --no-cpu-baseline --no-dense-probe > gpurun_out/r2y_stream1.json 2> gpurun_out/r2y_stream1.err
python - <<'PY'
import json
for f in ('r2y_rot','r2y_stream1'):
    try:
        j=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f, 'value',round(j['value'],1),'ms',round(j['ms_per_step'],4),'e2e',round(j['e2e']['value'],1),'seq',round(j['e2e']['sequential_value'],1),'GN us/pass',round(j['roofline']['us_per_launch'],2),'step_ms',j.get('step_ms',{}).get('resident'))
        print('    breakdown', j.get('step_breakdown_ms'), 'pose_err', j.get('pose_err_m'))
    except Exception as e: print(f,'ERR',e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
PY
