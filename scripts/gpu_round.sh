#!/bin/bash
# One GPU-box session: parity tests, smoke, bench lines, ncu launch list + full capture of the sweep kernel.
# Usage (under gpurun): bash scripts/gpu_round.sh <label> [stages...]   stages: test smoke bench benchab ncu
set -u
LABEL=${1:-run}; shift || true
STAGES=${*:-"test smoke bench benchab ncu"}
OUT=gpurun_out/$LABEL
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.csv" 2>&1
for s in $STAGES; do
  case $s in
    test)   timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/rc.log"; tail -5 "$OUT/pytest_gpu.log";;
    smoke)  timeout 300 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/rc.log"; tail -3 "$OUT/smoke.log";;
    bench)  timeout 900 python bench.py --steps 3 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/rc.log"; cat "$OUT/bench.json"; tail -3 "$OUT/bench.err";;
    benchab) timeout 600 python bench.py --steps 3 --warmup 3 --flags 1 --no-e2e --no-cpu > "$OUT/bench_ab.json" 2> "$OUT/bench_ab.err"; echo "benchab rc=$?" | tee -a "$OUT/rc.log"; cat "$OUT/bench_ab.json";;
    quick)  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > "$OUT/bench_quick.json" 2> "$OUT/bench_quick.err"; echo "quick rc=$?" | tee -a "$OUT/rc.log"; python -c "
import json,sys
d=json.loads(open('$OUT/bench_quick.json').read().strip().splitlines()[-1]); print('QUICK value=%.1f GLUPS frac=%.3f variant=%s 512^3=%.1f clocks=%s' % (d['value'], d['roofline']['frac'], d['sweep_variant'], d['extra'].get('glups_512cubed_1gpu',0), d['clocks']))"; tail -2 "$OUT/bench_quick.err";;
    tbtest) timeout 900 python -m pytest tests -m gpu -x -q -k "time_tiled or division or sweep_bitwise or default_wafer or simple_sweep" > "$OUT/pytest_tb.log" 2>&1; echo "tbtest rc=$?" | tee -a "$OUT/rc.log"; tail -3 "$OUT/pytest_tb.log";;
    ref)    timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "ref rc=$?" | tee -a "$OUT/rc.log"; cat "$OUT/bench_ref.json";;
    ncu2)
      # launch list of a short bench step at 512^3 (shares of the step) + full captures of every hot kernel.  Only the
      # CSV exports travel back (gpurun_out is capped at 64 MiB); of the reports just the time-tiled sweep's is kept.
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" \
        python bench.py --grid 512 --sweeps 100 --steps 2 --warmup 1 --no-cpu --no-512 --no-parity > "$OUT/ncu_launch_bench.log" 2>&1
      echo "ncu-launches rc=$?" | tee -a "$OUT/rc.log"
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep_tb2" -s 20 -c 1 -f -o "$OUT/tb2_full" \
        python scripts/ncu_workload.py > "$OUT/ncu_tb2.log" 2>&1
      echo "ncu-tb2 rc=$?" | tee -a "$OUT/rc.log"
      ncu -i "$OUT/tb2_full.ncu-rep" --page raw --csv > "$OUT/tb2_full_raw.csv" 2>/dev/null
      timeout 1200 ncu --set full --clock-control none -k regex:"sweep_tma1|project_kernel|dots_kernel|gs_coeff|finalize" -f -o /tmp/others_full \
        python scripts/ncu_workload.py > "$OUT/ncu_others.log" 2>&1
      echo "ncu-others rc=$?" | tee -a "$OUT/rc.log"
      ncu -i /tmp/others_full.ncu-rep --page raw --csv > "$OUT/others_full_raw.csv" 2>/dev/null
      ls -la "$OUT" /tmp/others_full.ncu-rep;;
    sanitize)
      timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_small.py > "$OUT/memcheck.log" 2>&1; echo "memcheck rc=$?" | tee -a "$OUT/rc.log"; tail -2 "$OUT/memcheck.log"
      timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_small.py > "$OUT/racecheck.log" 2>&1; echo "racecheck rc=$?" | tee -a "$OUT/rc.log"; tail -2 "$OUT/racecheck.log";;
    ncu_tb2)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" \
        python bench.py --grid 512 --sweeps 100 --steps 2 --warmup 1 --no-cpu --no-512 --no-parity > "$OUT/ncu_launch_bench.log" 2>&1
      echo "ncu-launches rc=$?" | tee -a "$OUT/rc.log"
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep_tb2" -s 20 -c 1 -f -o "$OUT/tb2_full" \
        python scripts/ncu_workload.py > "$OUT/ncu_tb2.log" 2>&1
      echo "ncu-tb2 rc=$?" | tee -a "$OUT/rc.log"
      ncu -i "$OUT/tb2_full.ncu-rep" --page raw --csv > "$OUT/tb2_full_raw.csv" 2>/dev/null;;
    c3)     WAFER_SLOW_TESTS=1 timeout 1500 python scripts/config_parity.py C3 --steps 50 --screen 50 > "$OUT/config_parity_C3.json" 2> "$OUT/config_parity_C3.err"; echo "c3 rc=$?" | tee -a "$OUT/rc.log"; tail -30 "$OUT/config_parity_C3.json";;
    c2)     timeout 900 python scripts/config_parity.py C2 --steps 50 --screen 50 > "$OUT/config_parity_C2.json" 2> "$OUT/config_parity_C2.err"; echo "c2 rc=$?" | tee -a "$OUT/rc.log"; tail -30 "$OUT/config_parity_C2.json";;
    extra)  N=512 M=512 timeout 900 python scripts/extra_bench.py > "$OUT/extra.json" 2> "$OUT/extra.err"; echo "extra rc=$?" | tee -a "$OUT/rc.log"; cat "$OUT/extra.json"; tail -3 "$OUT/extra.err";;
    ncu)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
        python bench.py --grid 512 --sweeps 20 --steps 2 --warmup 1 --no-e2e --no-cpu --no-512 > "$OUT/ncu_launch_bench.log" 2>&1
      echo "ncu-launches rc=$?" | tee -a "$OUT/rc.log"
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep -s 6 -c 2 -f -o "$OUT/sweep_full" \
        python bench.py --grid 512 --sweeps 20 --steps 2 --warmup 1 --no-e2e --no-cpu --no-512 > "$OUT/ncu_full_bench.log" 2>&1
      echo "ncu-full rc=$?" | tee -a "$OUT/rc.log";;
  esac
done
