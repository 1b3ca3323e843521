#!/bin/bash
# Diagnostic (GPU box): schedule order of the pair list -- keyframes per schedule group (0 = frame-major over all keyframes).
for g in 0 1 2 4 8 16; do
  ELLC_ORDER_KF_GROUP=$g timeout 200 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('kf_group=$g', round(d['value']), 'kernel ms', round(d['roofline']['kernel_ms_per_launch'],2))"
done
