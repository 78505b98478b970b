#!/bin/bash
# Multi-GPU session (gpurun --gpus N): slab parity check, then the bench scaling series on 1..N GPUs.
set -u
LABEL=${1:-multi}; N=${2:-2}; shift 2 || true
EXTRA=${*:-}
OUT=gpurun_out/$LABEL
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
for mode in ${CHECK_MODES:-1 0}; do
  WAFER_P2P=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29601+mode)) \
      scripts/multigpu_check.py > "$OUT/multigpu_check_p2p$mode.log" 2>&1
  echo "multigpu_check p2p=$mode rc=$?" | tee -a "$OUT/rc.log"; grep '^{' "$OUT/multigpu_check_p2p$mode.log" | tail -1; tail -2 "$OUT/multigpu_check_p2p$mode.log" | cut -c1-300
done
for n in ${NLIST:-1 2 4 8}; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu --no-512 $EXTRA > "$OUT/scale_$n.json" 2> "$OUT/scale_$n.err"
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) \
      bench.py --gpus $n --steps 3 --warmup 3 $EXTRA > "$OUT/scale_$n.json" 2> "$OUT/scale_$n.err"
  fi
  echo "scale $n rc=$?" | tee -a "$OUT/rc.log"; grep '^{' "$OUT/scale_$n.json" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('N=%d value=%.1f GLUPS e2e=%s frac=%.3f' % (d['n_gpus'], d['value'], d['e2e'] and round(d['e2e']['value'],1), d['roofline']['frac']))"
  tail -2 "$OUT/scale_$n.err"
done
