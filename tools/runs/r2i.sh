#!/bin/bash
# round-2 GPU run I (1 GPU): lanes-per-query sweep for large query sets after the pruning changes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for L in 1 2 4 8; do
  for w in "10000000 hdl" "1000000 x8" "1000000 dense"; do
    echo "lanes$L $w: $(LILIOM_KNN_LANES=$L timeout 200 python tools/knn_once.py $w 2>&1 | tail -1)" >> gpurun_out/r2i_lanes.log
  done
done
for R in 2 4; do
  echo "lanes4 rounds$R 10000000 hdl: $(LILIOM_KNN_LANES=4 LILIOM_KNN_ROUNDS=$R timeout 200 python tools/knn_once.py 10000000 hdl 2>&1 | tail -1)" >> gpurun_out/r2i_lanes.log
done
cat gpurun_out/r2i_lanes.log
