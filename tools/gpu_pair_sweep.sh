#!/bin/bash
# BASELINE config 5 at one GPU: frame-pair throughput vs batch size (pairs = 9 x frames; frames and keyframes resident).
for fk in "28 8" "114 16" "512 32" "1024 32" "2048 64" "4096 64"; do
  set -- $fk
  timeout 900 python bench.py --frames $1 --keyframes $2 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; c=d['config']
print('frames %5d keyframes %3d pairs %6d : %8.0f tracks/s  step %7.2f ms  kernel %7.2f ms  frac %.3f' % (c['frames_per_gpu'], c['keyframes_per_gpu'], c['pairs_per_gpu_per_step'], d['value'], d['ms_per_step'], r['kernel_ms_per_launch'], r['frac']))"
done
