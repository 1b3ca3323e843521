#!/bin/bash
# Diagnostic (GPU box): bench every CTA-size build variant under build_variants/ (kernel-only numbers).
for so in build_variants/*.so; do
  ELLC_LIB=$PWD/$so timeout 200 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$so', round(d['value']), round(d['roofline']['kernel_ms_per_launch'],2))"
done
