#!/bin/bash
# BASELINE configs C2 and C3 through the front end on one GPU: bash scripts/gpu_configs.sh <label> [c2] [c3]
set -u
LABEL=$1; shift
OUT=gpurun_out/$LABEL; mkdir -p "$OUT"
for c in "$@"; do
  case $c in
    c2) cfg=examples/c2_harmonic_256.yaml; lim=300;;
    c3) cfg=examples/c3_cornell_512.yaml; lim=${C3_LIMIT:-900};;
  esac
  mkdir -p "$OUT/$c"
  t0=$(date +%s.%N)
  timeout $lim wafer_b200/wafer-b200 -c $cfg --output-root "$OUT/$c/out" > "$OUT/$c/run.log" 2> "$OUT/$c/run.err"
  rc=$?
  echo "$c rc=$rc wall=$(python -c "import time; print('%.1f s' % (time.time() - $t0))")" | tee "$OUT/$c/wall.txt"
  find "$OUT/$c/out" -name 'observables_*' -exec cp {} "$OUT/$c/" \;
  rm -rf "$OUT/$c/out"
  grep -E "energy =" "$OUT/$c/run.log"
done
