#!/bin/bash
# Round 2: after the ring-entry fix (all idle entries grow together): short default runs must no longer stall at the first timed step; final
# bench lines, launch list and full capture of the final build.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/q_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/q_tests.log
tail -4 gpurun_out/q_tests.log
for r in 1 2 3; do
  timeout 600 python bench.py --no-cpu-baseline > gpurun_out/q_default_$r.json 2> gpurun_out/q_default_$r.err
  echo "default $r rc=$? $(python -c "import json;j=json.load(open('gpurun_out/q_default_$r.json'));h=j.get('host_ms_per_step',{});print(round(j['value']), round(j['ms_per_step'],3), round(j['roofline']['kernel_ms_per_launch'],3), 'e2e', round(j['e2e']['value']), {k:(round(v['median'],2),round(v['max'],1),v['argmax_step']) for k,v in h.items() if isinstance(v,dict)})" 2>&1 | tail -1)"
done
for r in 1 2; do
  timeout 600 python bench.py --lc-mode const_weight --steps 10 --warmup 3 > gpurun_out/q_lc_$r.json 2> gpurun_out/q_lc_$r.err
  echo "lc $r rc=$? $(python -c "import json;j=json.load(open('gpurun_out/q_lc_$r.json'));h=j.get('host_ms_per_step',{});print(round(j['value']), round(j['ms_per_step'],3), round(j['roofline']['kernel_ms_per_launch'],3), {k:(round(v['median'],2),round(v['max'],1),v['argmax_step']) for k,v in h.items() if isinstance(v,dict)})" 2>&1 | tail -1)"
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/q_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/q_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_track -s 2 -c 1 -f -o gpurun_out/prof_r2_q python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/q_ncu_full.log 2>&1; echo "ncu full rc=$?"
