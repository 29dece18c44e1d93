#!/bin/bash
# round-2 GPU run U (1 GPU): k_rot_ring warp walk, block-local boxes, block-reduced k_vg_minmax; stream launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2u_tests.log; tail -3 gpurun_out/r2u_tests.log
timeout 600 python bench.py --workload stream --steps 20 --warmup 3 --no-cpu-baseline --no-dense-probe > gpurun_out/r2u_stream1.json 2> gpurun_out/r2u_stream1.err
timeout 600 python bench.py --workload rot --steps 100 --warmup 5 --no-cpu-baseline --no-dense-probe > gpurun_out/r2u_rot.json 2> gpurun_out/r2u_rot.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2u_stream_launches.csv python bench.py --workload stream --steps 2 --warmup 2 --no-cpu-baseline --no-dense-probe --e2e sequential --no-extra-legs > gpurun_out/r2u_ncu_stream.log 2>&1
python - <<'PY'
import json
for f in ('r2u_stream1','r2u_rot'):
    try:
        j=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f, 'value',round(j['value'],1),'ms',round(j['ms_per_step'],4),'e2e',round(j['e2e']['value'],1),'seq',round(j['e2e']['sequential_value'],1),'step_ms',j.get('step_ms',{}).get('resident'))
        print('    breakdown', j.get('step_breakdown_ms'), 'inc', (j.get('incremental_map') or {}).get('value'), (j.get('incremental_map') or {}).get('map_update_ms'), 'pose_err', j.get('pose_err_m'))
    except Exception as e: print(f,'ERR',e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
PY
