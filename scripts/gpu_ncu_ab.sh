#!/bin/bash
# ncu --set full of the time-tiled sweep under two schedules at one size: bash scripts/gpu_ncu_ab.sh <label> <grid>
set -u
LABEL=$1; GRID=${2:-1024}
OUT=gpurun_out/$LABEL; mkdir -p "$OUT"
for sched in rounds legacy; do
  if [ $sched = legacy ]; then export WAFER_TB_SCHED=legacy; else unset WAFER_TB_SCHED; fi
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_tb2 -s 6 -c 2 -f -o "$OUT/tb2_${GRID}_$sched" \
    python bench.py --grid $GRID --sweeps 20 --steps 2 --warmup 1 --no-e2e --no-cpu --no-512 --no-parity > "$OUT/ncu_$sched.log" 2>&1
  echo "ncu $sched rc=$?"
done
