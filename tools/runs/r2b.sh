#!/bin/bash
# round-2 GPU run B (2 GPUs): sharded parity (NCCL and fused peer exchange), then the N=2 bench line of the streamed workload
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multi_check.py 1000000 > gpurun_out/r2b_check_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_check_nccl.log
LILIOM_PEER=1 timeout 300 $TR tools/multi_check.py 1000000 > gpurun_out/r2b_check_peer.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_check_peer.log
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --map-points 2000000 > gpurun_out/r2b_bench2_small.json 2> gpurun_out/r2b_bench2_small.err; echo "rc=$?" >> gpurun_out/r2b_bench2_small.err
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2b_bench2.json 2> gpurun_out/r2b_bench2.err; echo "rc=$?" >> gpurun_out/r2b_bench2.err
tail -4 gpurun_out/r2b_check_nccl.log gpurun_out/r2b_check_peer.log; tail -3 gpurun_out/r2b_bench2_small.err; cat gpurun_out/r2b_bench2_small.json; tail -3 gpurun_out/r2b_bench2.err; cat gpurun_out/r2b_bench2.json
