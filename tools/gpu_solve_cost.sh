#!/bin/bash
# Diagnostic (GPU box): what does the serial K5 solve cost?  Same pixel work (fixed iteration counts, no early-out), with
# and without the solve, for 1 / 2 / 4 pairs per CTA.
export ELLC_DEBUG_MAX_ITER=3,3,4,5 ELLC_DEBUG_NO_STOP=1
for nu in 0 1; do for np in 1 2 4; do
  ELLC_DEBUG_NO_UPDATE=$nu ELLC_PAIRS_PER_CTA=$np timeout 200 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('no_update=$nu np=$np', round(d['value']), 'kernel ms', round(r['kernel_ms_per_launch'],2), r['mean_iters_per_level'])"
done; done
