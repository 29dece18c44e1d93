#!/bin/bash
# round-2 GPU run V (1 GPU): grid box escape path, stream bench, ncu source profile of k_rot_ring
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2v_tests.log; tail -3 gpurun_out/r2v_tests.log
timeout 600 python bench.py --workload stream --steps 20 --warmup 3 --no-cpu-baseline --no-dense-probe > gpurun_out/r2v_stream1.json 2> gpurun_out/r2v_stream1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rot_ring -s 3 -c 1 -f -o gpurun_out/r2v_rot_ring python bench.py --workload rot --steps 2 --warmup 2 --no-cpu-baseline --no-dense-probe --e2e sequential > gpurun_out/r2v_ncu_ring.log 2>&1
python - <<'PY'
import json
for f in ('r2v_stream1',):
    try:
        j=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f, 'value',round(j['value'],1),'ms',round(j['ms_per_step'],4),'e2e',round(j['e2e']['value'],1),'seq',round(j['e2e']['sequential_value'],1),'step_ms',j.get('step_ms',{}).get('resident'))
        print('    breakdown', j.get('step_breakdown_ms'), 'inc', (j.get('incremental_map') or {}).get('value'), (j.get('incremental_map') or {}).get('map_update_ms'), 'pose_err', j.get('pose_err_m'))
    except Exception as e: print(f,'ERR',e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
PY
