#!/bin/bash
# Round 2, closing pass on one GPU (final build: batch overlap on by default): bench lines, launch list, full capture.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "bench rc=$?"
timeout 600 python bench.py > gpurun_out/u_bench_default.json 2> gpurun_out/u_bench_default.err; echo "bench default rc=$?"
timeout 600 python bench.py --lc-mode const_weight --steps 20 --warmup 3 > gpurun_out/u_bench_lc.json 2> gpurun_out/u_bench_lc.err; echo "lc rc=$?"
for f in u_bench u_bench_default u_bench_lc; do python -c "import json;j=json.load(open('gpurun_out/$f.json'));print('$f', round(j['value']), round(j['ms_per_step'],3), round(j['roofline']['kernel_ms_per_launch'],3), 'e2e', j['e2e'] and round(j['e2e']['value']), j['host_ms_per_step']['launch']['max'])"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/u_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/u_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_track -s 2 -c 1 -f -o gpurun_out/prof_r2_u python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/u_ncu_full.log 2>&1; echo "ncu full rc=$?"
