#!/bin/bash
# 2-GPU A/B of the fused-halo forms at the slab size of an 8-GPU run (128 planes per rank): whole-column (default) vs
# split boundary launches (WAFER_P2P_SPLIT=1), after the parity checks of the new default.
set -u
LABEL=${1:-multi2b}; N=2
OUT=gpurun_out/$LABEL; mkdir -p "$OUT"
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
WAFER_P2P=1 tr 29601 scripts/multigpu_check.py > "$OUT/check_p2p.log" 2>&1; echo "check rc=$?" | tee -a "$OUT/rc.log"; grep '^{' "$OUT/check_p2p.log" | tail -1
WAFER_P2P=1 WAFER_DEBUG_SKIP_FINAL_WAIT=1 tr 29602 scripts/multigpu_check.py > "$OUT/check_p2p_neg.log" 2>&1; echo "check-negative rc=$? (must be 1)" | tee -a "$OUT/rc.log"; grep '^{' "$OUT/check_p2p_neg.log" | tail -1 | cut -c1-80
for mode in 0 1; do
  WAFER_P2P_SPLIT=$mode tr $((29610+mode)) bench.py --gpus 2 --nx 256 --steps 3 --warmup 3 --no-e2e --no-512 > "$OUT/slab128_split$mode.json" 2> "$OUT/slab128_split$mode.err"; echo "slab128 split=$mode rc=$?" | tee -a "$OUT/rc.log"
  WAFER_P2P_SPLIT=$mode tr $((29620+mode)) bench.py --gpus 2 --steps 3 --warmup 3 > "$OUT/c4_split$mode.json" 2> "$OUT/c4_split$mode.err"; echo "c4 split=$mode rc=$?" | tee -a "$OUT/rc.log"
done
timeout 600 python bench.py --gpus 1 --nx 128 --steps 3 --warmup 3 --no-cpu --no-512 --no-e2e --no-parity > "$OUT/slab128_1gpu.json" 2> "$OUT/slab128_1gpu.err"
for f in slab128_split0 slab128_split1 slab128_1gpu c4_split0 c4_split1; do grep '^{' "$OUT/$f.json" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); e = d.get('e2e') or {}; p = d.get('parity') or {}
    print('$f N=%d value=%.1f GLUPS e2e=%s E=%s parity=%s clocks=%s' % (d['n_gpus'], d['value'], e.get('value'), e.get('last_energy'), (p.get('ok'), p.get('energy_rel_diff')), d['clocks']['sm_mhz']))"; done
