#!/bin/bash
# round-2 GPU run G (1 GPU): search and fit as two kernels for large query sets, A/B against the fused kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2g_tests.log
for cfg in "split64:" "fused:LILIOM_KNN_FUSED=1" "split80:LILIOM_LIB=$PWD/liliom_b200/libliliom_b200_s3.so"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  for w in "10000000 hdl" "1000000 hdl" "1000000 x8"; do
    echo "$name $w: $(env $envs timeout 200 python tools/knn_once.py $w 2>&1 | tail -1)" >> gpurun_out/r2g_dense.log
  done
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_knn_search1 -s 12 -c 1 -f -o gpurun_out/r2g_search python tools/knn_once.py 10000000 hdl > gpurun_out/r2g_ncu.log 2>&1
tail -4 gpurun_out/r2g_tests.log; cat gpurun_out/r2g_dense.log
