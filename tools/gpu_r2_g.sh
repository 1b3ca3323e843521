#!/bin/bash
# Round 2, instruction diet of the forward kernel, second pass: parity tests, A/B against the previous build on the same box (forward and
# loop-closure mode), bench line, ncu launch list + full capture.
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g_tests.log
tail -15 gpurun_out/g_tests.log
for v in base old lane; do
  lib=build/variants/libellc_gn_$v.so
  [ $v = base ] && lib=egomotion_with_local_loop_closures_b200/libellc_gn.so
  ELLC_LIB=$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/g_var_$v.json 2> gpurun_out/g_var_$v.err
  echo "$v rc=$? $(python -c "import json;j=json.load(open('gpurun_out/g_var_$v.json'));print(round(j['value']), j['roofline']['kernel_ms_per_launch'])" 2>&1 | tail -1)"
done
for v in base old; do
  lib=build/variants/libellc_gn_$v.so
  [ $v = base ] && lib=egomotion_with_local_loop_closures_b200/libellc_gn.so
  ELLC_LIB=$lib timeout 600 python bench.py --lc-mode const_weight --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/g_lc_$v.json 2> gpurun_out/g_lc_$v.err
  echo "lc $v rc=$? $(python -c "import json;j=json.load(open('gpurun_out/g_lc_$v.json'));print(round(j['value']), j['roofline']['kernel_ms_per_launch'])" 2>&1 | tail -1)"
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"
head -c 600 gpurun_out/g_bench.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/g_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_track -s 2 -c 1 -f -o gpurun_out/prof_r2_g python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/g_ncu_full.log 2>&1; echo "ncu full rc=$?"
