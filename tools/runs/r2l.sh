#!/bin/bash
# round-2 GPU run L (1 GPU): bulk-copy (TMA) staging variant of the 16-lane search: bit identity + A/B; SASS evidence
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2l_tests.log
timeout 400 python tools/ab_variants.py 150 3 33 3 33 > gpurun_out/r2l_ab.log 2>&1
LILIOM_KNN_TMA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gn_persistent -s 2 -c 1 -f -o gpurun_out/r2l_gn_tma python tools/knn_once.py 1000000 ds > gpurun_out/r2l_ncu.log 2>&1
tail -3 gpurun_out/r2l_tests.log; cat gpurun_out/r2l_ab.log
