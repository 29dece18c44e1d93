#!/bin/bash
# round-2 GPU run T (1 GPU): rebuild with host-known boxes + two-node e2e on the streamed workload; sampled kernel timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2t_tests.log; tail -3 gpurun_out/r2t_tests.log
LILIOM_DEBUG_TIMING=1 timeout 300 python bench.py --workload stream --steps 4 --warmup 3 --no-cpu-baseline --no-dense-probe --e2e sequential > gpurun_out/r2t_dbg_stream.json 2> gpurun_out/r2t_dbg_stream.err; grep "map_rebuild" gpurun_out/r2t_dbg_stream.err | tail -2
timeout 600 python bench.py --workload stream --steps 20 --warmup 3 --no-cpu-baseline --no-dense-probe > gpurun_out/r2t_stream1.json 2> gpurun_out/r2t_stream1.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-dense-probe > gpurun_out/r2t_bench20.json 2> gpurun_out/r2t_bench20.err
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline --no-dense-probe > gpurun_out/r2t_bench200.json 2> gpurun_out/r2t_bench200.err
python - <<'PY'
import json
for f in ('r2t_stream1','r2t_bench20','r2t_bench200'):
    try:
        j=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f, 'value',round(j['value'],1),'ms',round(j['ms_per_step'],4),'e2e',round(j['e2e']['value'],1),'seq',round(j['e2e']['sequential_value'],1),'step_ms',j.get('step_ms',{}).get('resident'),'roof',round(j['roofline']['us_per_launch'],2), j['roofline']['passes_timed'])
        print('    breakdown', j.get('step_breakdown_ms'), 'inc', (j.get('incremental_map') or {}).get('value'), (j.get('incremental_map') or {}).get('map_update_ms'), 'pose_err', j.get('pose_err_m'))
    except Exception as e: print(f,'ERR',e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
PY
