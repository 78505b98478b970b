#!/bin/bash
# A/B runs of prebuilt library variants on one GPU box: bash scripts/gpu_libs.sh <label> <name1> <name2> ...
# (names of build/variants/<name>.so; "default" = wafer_b200/libwafer_b200.so)
set -u
LABEL=$1; shift
OUT=gpurun_out/$LABEL; mkdir -p "$OUT"
for name in "$@"; do
  unset WAFER_TB_SCHED WAFER_TB_MAXSEG
  case $name in maxseg*) export WAFER_TB_MAXSEG=${name#maxseg}; name_lib=default;; *) name_lib=$name;; esac
  case $name in sched_*) export WAFER_TB_SCHED=${name#sched_}; name_lib=default;; esac
  if [ "$name_lib" = default ]; then unset WAFER_B200_LIB; elif [ "$name" = legacy ]; then unset WAFER_B200_LIB; export WAFER_TB_SCHED=legacy; else export WAFER_B200_LIB=$PWD/build/variants/$name.so; fi
  timeout 300 python -m pytest tests -m gpu -x -q -k "${TESTS:-time_tiled}" > "$OUT/pytest_$name.log" 2>&1; echo "$name tests rc=$? $(tail -1 $OUT/pytest_$name.log)"
  timeout 600 python bench.py --steps ${STEPS:-2} --warmup 2 --no-e2e --no-cpu --no-parity ${BENCH_ARGS:-} > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python -c "
import json
d=json.loads(open('$OUT/bench_$name.json').read().strip().splitlines()[-1]); print('$name value=%.1f GLUPS 512^3=%.1f clocks=%s' % (d['value'], d['extra'].get('glups_512cubed_1gpu',0), d['clocks']['sm_mhz']))" || tail -3 "$OUT/bench_$name.err"
done
