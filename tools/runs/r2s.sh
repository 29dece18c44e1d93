#!/bin/bash
# round-2 GPU run S (1 GPU): zero-copy results A/B, step timeline, 20-step bench with the looping clock sampler
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s_tests.log; tail -3 gpurun_out/r2s_tests.log
timeout 300 python tools/ab_variants.py 150 3 43 103 3 43 103 > gpurun_out/r2s_ab.log 2>&1; cat gpurun_out/r2s_ab.log
LILIOM_DEBUG_TIMING=1 timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-dense-probe --e2e sequential > gpurun_out/r2s_dbg.json 2> gpurun_out/r2s_dbg.err; grep "timeline" gpurun_out/r2s_dbg.err | tail -3
for i in 1 2 3; do
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-dense-probe > gpurun_out/r2s_bench20_$i.json 2> gpurun_out/r2s_bench20_$i.err
done
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline --no-dense-probe > gpurun_out/r2s_bench200.json 2> gpurun_out/r2s_bench200.err
python - <<'PY'
import json
for f in ('r2s_bench20_1','r2s_bench20_2','r2s_bench20_3','r2s_bench200'):
    try:
        j=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f, 'value',round(j['value'],1),'e2e',round(j['e2e']['value'],1),'seq',round(j['e2e']['sequential_value'],1),'step_ms',j.get('step_ms'),'clocks',j['clocks'])
    except Exception as e: print(f,'ERR',e)
PY
