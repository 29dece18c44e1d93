#!/bin/bash
# round-2 last GPU run (1 GPU): the final commit — full GPU test tier, smoke, the driver's bench command
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2fin_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2fin_tests.log; tail -3 gpurun_out/r2fin_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2fin_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2fin_smoke.log; tail -2 gpurun_out/r2fin_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2fin_bench.json 2> gpurun_out/r2fin_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2fin_bench.json') if l.startswith('{')][-1])
print('value',round(j['value'],1),'e2e',round(j['e2e']['value'],1),'seq',round(j['e2e']['sequential_value'],1),'cpu',j['cpu_baseline']['value'],'roof',j['roofline']['frac'],'clocks',j['clocks'],'launches',j['gpu_launches'])
PY
