#!/bin/bash
# round-2 GPU run F (1 GPU): temporal-coherence bound in the search — suite, dense timings, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log
for w in "1000000 hdl" "10000000 hdl" "1000000 x8" "1000000 dense" "1000000 ds"; do
  echo "$w: $(timeout 200 python tools/knn_once.py $w 2>&1 | tail -1)" >> gpurun_out/r2f_dense.log
done
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -4 gpurun_out/r2f_tests.log; cat gpurun_out/r2f_dense.log; cut -c1-300 gpurun_out/r2f_bench.json; python - <<'PY'
import json
j=json.load(open('gpurun_out/r2f_bench.json'))
r=j['roofline']
print({k:r[k] for k in ('queries_per_launch','candidates_per_query','examined_per_query','us_per_launch','frac')})
for k,v in r.get('dense_probe',{}).items(): print(k,{kk:v.get(kk) for kk in ('queries_per_launch','candidates_per_query','examined_per_query','us_per_launch','frac','error')})
print('e2e',j['e2e']['value'],j['e2e']['sequential_value'])
PY
