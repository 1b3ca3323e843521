#!/bin/bash
# Round 2: schedule-shape experiments on the slimmer kernel (fewer concurrent frames => fewer L2 misses on the texel gathers?)
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/h_$name.json 2> gpurun_out/h_$name.err
  echo "$name rc=$? $(python -c "import json;j=json.load(open('gpurun_out/h_$name.json'));print(round(j['value']), j['roofline']['kernel_ms_per_launch'])" 2>&1 | tail -1)"
}
run base
run cluster2 --cluster 2
run cluster4 --cluster 4
run np2 --pairs-per-cta 2
run np4 --pairs-per-cta 4
run base_lc --lc-mode const_weight
