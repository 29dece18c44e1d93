#!/bin/bash
# A/B of the small-scan GN kernel variants on one B200 (run under gpurun): parity first, then bench lines.
# usage: tools/ab_gn.sh [steps]
STEPS=${1:-200}
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests -m gpu -x -q ) > gpurun_out/ab_tests_default.log 2>&1; tail -4 gpurun_out/ab_tests_default.log
( LILIOM_KNN_FLAT=2 LILIOM_GN_LL=1 timeout 200 python -m pytest tests -m gpu -x -q ) > gpurun_out/ab_tests_flat2_ll1.log 2>&1; tail -3 gpurun_out/ab_tests_flat2_ll1.log
for v in "0 0" "1 0" "2 0" "0 1" "2 1"; do
  set -- $v
  LILIOM_KNN_FLAT=$1 LILIOM_GN_LL=$2 timeout 100 python bench.py --steps $STEPS --no-cpu-baseline > gpurun_out/ab_flat$1_ll$2.json 2> gpurun_out/ab_flat$1_ll$2.err
  python - "$1" "$2" <<'PY'
import json,sys
f,l=sys.argv[1:3]
try:
    d=json.load(open(f"gpurun_out/ab_flat{f}_ll{l}.json"))
    print(f"flat={f} ll={l}: value {d['value']:.0f} scans/s  e2e {d['e2e']['value']:.0f} (seq {d['e2e']['sequential_value']:.0f})  knn {d['roofline']['us_per_launch']:.2f} us/pass  q/launch {d['roofline']['queries_per_launch']:.0f}  pose_err {d['pose_err_m']:.5f}")
except Exception as e:
    print(f"flat={f} ll={l}: FAILED {e}")
PY
done
LILIOM_KNN_FLAT=2 LILIOM_GN_LL=1 LILIOM_DEBUG_TIMING=1 timeout 60 python tools/knn_once.py 1000000 ds 2>&1 | grep -E "persistent|us/launch" | tail -3
