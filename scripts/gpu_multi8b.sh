#!/bin/bash
# short 8-GPU session for the whole-column fused halo: parity check, C4 bench at N = 8 and its N = 1 point
set -u
LABEL=${1:-multi8b}; N=8
OUT=gpurun_out/$LABEL; mkdir -p "$OUT"
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
WAFER_P2P=1 tr 29601 scripts/multigpu_check.py > "$OUT/check_p2p.log" 2>&1; echo "check rc=$?" | tee -a "$OUT/rc.log"; grep '^{' "$OUT/check_p2p.log" | tail -1 | cut -c1-400
tr 29603 bench.py --gpus $N --steps 3 --warmup 3 > "$OUT/scale_$N.json" 2> "$OUT/scale_$N.err"; echo "bench$N rc=$?" | tee -a "$OUT/rc.log"
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu --no-512 --no-e2e --no-parity > "$OUT/scale_1.json" 2> "$OUT/scale_1.err"; echo "bench1 rc=$?" | tee -a "$OUT/rc.log"
for f in scale_$N scale_1; do grep '^{' "$OUT/$f.json" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); e = d.get('e2e') or {}; p = d.get('parity') or {}
    print('$f N=%d value=%.1f GLUPS e2e=%s E=%s parity=%s clocks=%s' % (d['n_gpus'], d['value'], e.get('value'), e.get('last_energy'), (p.get('ok'), p.get('energy_rel_diff')), d['clocks']['sm_mhz']))"; done
