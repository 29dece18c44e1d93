#!/bin/bash
# round-2 GPU run Z (1 GPU): final evidence — smoke, the driver's two bench arms, launch list, ncu full capture of the GN kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2z_tests.log; tail -3 gpurun_out/r2z_tests.log
LILIOM_DEBUG_TIMING=1 timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-dense-probe --e2e sequential > gpurun_out/r2z_dbg.json 2> gpurun_out/r2z_dbg.err; grep "coop" gpurun_out/r2z_dbg.err | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_smoke.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err ) 2> gpurun_out/r2z_time_reference.txt
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err ) 2> gpurun_out/r2z_time_bench.txt
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 --no-dense-probe > gpurun_out/r2z_bench200.json 2> gpurun_out/r2z_bench200.err
timeout 600 python bench.py --workload rot --steps 100 --warmup 5 --no-dense-probe > gpurun_out/r2z_bench_rot.json 2> gpurun_out/r2z_bench_rot.err
timeout 600 python bench.py --workload stream --steps 20 --warmup 5 --no-dense-probe > gpurun_out/r2z_bench_stream.json 2> gpurun_out/r2z_bench_stream.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2z_stream_launches.csv python bench.py --workload stream --steps 2 --warmup 2 --no-cpu-baseline --no-dense-probe --e2e sequential --no-extra-legs > gpurun_out/r2z_ncu_stream.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-dense-probe > gpurun_out/r2z_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gn_persistent -s 2 -c 1 -f -o gpurun_out/r2z_gn python tools/knn_once.py 1000000 ds > gpurun_out/r2z_ncu_gn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_knn_plane -s 22 -c 1 -f -o gpurun_out/r2z_dense python tools/knn_once.py 10000000 hdl > gpurun_out/r2z_ncu_dense.log 2>&1
cat gpurun_out/r2z_smoke.log; cat gpurun_out/r2z_time_reference.txt gpurun_out/r2z_time_bench.txt; cut -c1-600 gpurun_out/r2z_bench_reference.json; python - <<'PY'
import json
for f in ('r2z_bench','r2z_bench200','r2z_bench_rot','r2z_bench_stream'):
    j=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
    r=j['roofline']
    print(f, 'value',round(j['value'],1),'e2e',round(j['e2e']['value'],1),'seq',round(j['e2e']['sequential_value'],1),'cpu',(j.get('cpu_baseline') or {}).get('value'),'us/pass',round(r['us_per_launch'],2),'frac',round(r['frac'],4),'launches',j['gpu_launches'],'clocks',j['clocks'])
    for k,v in (r.get('dense_probe') or {}).items(): print('   ',k,{kk:v.get(kk) for kk in ('queries_per_launch','us_per_launch','frac','error')})
PY
