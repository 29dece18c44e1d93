#!/bin/bash
# round-2 GPU run X4 (4 GPUs): the sharded streamed bench line with the final code
N=${1:-4}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2x4_bench_$N.json 2> gpurun_out/r2x4_bench_$N.err; echo "rc=$?" >> gpurun_out/r2x4_bench_$N.err
tail -2 gpurun_out/r2x4_bench_$N.err; python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/r2x4_bench_$N.json") if l.startswith("{")][-1])
print('value',j['value'],'ms',j['ms_per_step'],'ranks',j.get('ms_per_step_ranks'),'pose_err',j['pose_err_m'])
s=j.get('same_workload_1gpu') or {}
print('same1',s.get('value'),s.get('pose_max_abs_diff_vs_sharded'),'speedup',j.get('speedup_vs_1gpu_same_workload'))
print('repl',(j.get('replicas') or {}).get('value'),'e2e',j['e2e']['value'],'note',j.get('note'))
print('roof us/pass',j['roofline']['us_per_launch'],'run',j['run']['map_points_installed'],'breakdown',j.get('step_breakdown_ms'))
PY
