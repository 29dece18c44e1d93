#!/bin/bash
# round-2 GPU run C (1 GPU): ncu captures of the dense (one thread per query) and headline GN kernels, switch A/B, stage clocks,
# and the streamed workload on one GPU with its launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_knn_plane -s 22 -c 2 -f -o gpurun_out/r2c_dense python tools/knn_once.py 10000000 hdl > gpurun_out/r2c_ncu_dense.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gn_persistent -s 2 -c 1 -f -o gpurun_out/r2c_gn python tools/knn_once.py 1000000 ds > gpurun_out/r2c_ncu_gn.log 2>&1
timeout 400 python tools/ab_variants.py 150 > gpurun_out/r2c_ab.log 2>&1
LILIOM_DEBUG_TIMING=1 timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-dense-probe --e2e sequential > gpurun_out/r2c_dbg.json 2> gpurun_out/r2c_dbg.err
timeout 600 python bench.py --workload stream --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_stream1.json 2> gpurun_out/r2c_stream1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_stream_launches.csv python bench.py --workload stream --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_stream_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu_dense.log gpurun_out/r2c_ncu_gn.log; cat gpurun_out/r2c_ab.log; tail -12 gpurun_out/r2c_dbg.err; cat gpurun_out/r2c_stream1.json; tail -3 gpurun_out/r2c_stream1.err
