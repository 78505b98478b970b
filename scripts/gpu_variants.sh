#!/bin/bash
# Build-and-measure loop for kernel variants ON the GPU box (same image, nvcc present):
#   bash scripts/gpu_variants.sh <label> "<TBFLAGS variant 1>" "<TBFLAGS variant 2>" ...
set -u
LABEL=$1; shift
OUT=gpurun_out/$LABEL; mkdir -p "$OUT"
i=0
for flags in "$@"; do
  i=$((i+1))
  touch wafer_b200/csrc/sweep_tb.cuh
  make wafer_b200/libwafer_b200.so TBFLAGS="$flags" > "$OUT/build_$i.log" 2>&1 || { echo "variant $i build failed"; tail -5 "$OUT/build_$i.log"; continue; }
  grep -A2 "sweep_tb2" wafer_b200/csrc/ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo
  if [ "${MODE:-tb}" = "simple" ]; then
    touch wafer_b200/csrc/kernels.cuh
    grep -E "sweep_simple_kernelILi[123]ELb1ELb0" -A2 wafer_b200/csrc/ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo
    timeout 300 python -m pytest tests -m gpu -x -q -k "sweep_bitwise and not 256" > "$OUT/pytest_$i.log" 2>&1; echo "variant $i [$flags] tests rc=$? $(tail -1 $OUT/pytest_$i.log)"
    N=512 M=256 timeout 300 python scripts/extra_bench.py > "$OUT/extra_$i.json" 2> "$OUT/extra_$i.err"
    python -c "
import json
d=json.load(open('$OUT/extra_$i.json')); print('variant $i [$flags] ' + ' '.join('%s=%.1f' % (k.replace('_glups_512','').replace('_glups_256',''), v) for k, v in d.items() if 'glups' in k))"
  else
  timeout 300 python -m pytest tests -m gpu -x -q -k "time_tiled" > "$OUT/pytest_$i.log" 2>&1; echo "variant $i [$flags] tests rc=$? $(tail -1 $OUT/pytest_$i.log)"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > "$OUT/bench_$i.json" 2> "$OUT/bench_$i.err"
  python -c "
import json
d=json.loads(open('$OUT/bench_$i.json').read().strip().splitlines()[-1]); print('variant $i [$flags] value=%.1f GLUPS 512^3=%.1f clocks=%s' % (d['value'], d['extra'].get('glups_512cubed_1gpu',0), d['clocks']['sm_mhz']))"
  fi
done
