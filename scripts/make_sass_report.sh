#!/bin/bash
# SASS opcode histograms of every hot kernel of the built library -> profiles/r2_sass_hist.txt (no GPU needed)
set -u
LIB=${1:-wafer_b200/libwafer_b200.so}
OUT=${2:-profiles/r2_sass_hist.txt}
{
  echo "# SASS opcode histograms (cuobjdump -sass $LIB | scripts/sass_hist.py), whole kernel bodies."
  echo "# UTMALDG = cp.async.bulk.tensor (TMA) loads, SYNCS = mbarrier operations, DADD/DMUL/DFMA = the f64 arithmetic;"
  echo "# no HMMA/UTCMMA anywhere: the path is a stencil, tensor cores are not used."
  for k in sweep_tb2_kernelILb0 sweep_tb2_kernelILb1 "sweep_tma1_kernelILi1ELi0E" "sweep_tma1_kernelILi1ELi2E" "sweep_tma1_kernelILi1ELi9E" "sweep_tma1_kernelILi1ELi16E" \
           "sweep_tma1_kernelILi2ELi0E" "sweep_tma1_kernelILi3ELi0E" "project_kernelILi1ELb1" "dots_kernelILi1E" "gs_coeff_kernelILb1ELb1" "checksum_kernel" "sweep_simple_kernelILi1ELb1ELb0"; do
    python scripts/sass_hist.py "$LIB" "$k"
  done
  echo
  echo "# steady-state plane loops of the time-tiled sweep (two plane iterations per trip; cold out-of-line division blocks left out)"
  python scripts/sass_hist.py "$LIB" sweep_tb2_kernelILb0 --loops | grep -E "^   loop" | awk '{n=$4; gsub(/\(/,"",n); if (n+0 > 800 && n+0 < 1100) print}'
} > "$OUT" 2>&1
wc -l "$OUT"
