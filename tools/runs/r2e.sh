#!/bin/bash
# round-2 GPU run E (1 GPU): dense search v1b (own cell first) at 2 and 3 blocks per SM, suite, streamed workload after the
# map-maintenance fixes (frame buffers recycled, scan-based cell table)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e_tests.log
for v in "" _mb3; do
  for w in "1000000 hdl" "10000000 hdl" "1000000 x8" "1000000 dense"; do
    echo "lib$v $w: $(LILIOM_LIB=$PWD/liliom_b200/libliliom_b200$v.so timeout 200 python tools/knn_once.py $w 2>&1 | tail -1)" >> gpurun_out/r2e_dense.log
  done
done
timeout 600 python bench.py --workload stream --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_stream1.json 2> gpurun_out/r2e_stream1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2e_stream_launches.csv python bench.py --workload stream --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_stream_ncu.log 2>&1
tail -4 gpurun_out/r2e_tests.log; cat gpurun_out/r2e_dense.log; cut -c1-600 gpurun_out/r2e_stream1.json; tail -3 gpurun_out/r2e_stream1.err
