#!/bin/bash
# round-2 GPU run X (N GPUs, N = $1): sharded parity check with the fused exchange, then the streamed bench line
N=${1:-2}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
LILIOM_PEER=1 timeout 300 $TR tools/multi_check.py 1000000 > gpurun_out/r2x_check_peer_$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_check_peer_$N.log
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2x_bench_$N.json 2> gpurun_out/r2x_bench_$N.err; echo "rc=$?" >> gpurun_out/r2x_bench_$N.err
grep -v "^$" gpurun_out/r2x_check_peer_$N.log | tail -3 | cut -c1-300; tail -2 gpurun_out/r2x_bench_$N.err; python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/r2x_bench_$N.json") if l.startswith("{")][-1])
print('value',j['value'],'ms',j['ms_per_step'],'ranks',j.get('ms_per_step_ranks'),'pose_err',j['pose_err_m'])
print('same1',j.get('same_workload_1gpu'),'speedup',j.get('speedup_vs_1gpu_same_workload'))
print('repl',(j.get('replicas') or {}).get('value'),'e2e',j['e2e']['value'],'note',j.get('note'))
print('roof us/pass',j['roofline']['us_per_launch'],'run',j['run'],'breakdown',j.get('step_breakdown_ms'))
PY
