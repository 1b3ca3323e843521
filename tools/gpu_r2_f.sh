#!/bin/bash
# Round 2, instruction diet of the forward kernel: parity tests, A/B of the build variants on the same box, bench line, ncu launch list + full capture.
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/f_tests.log
tail -15 gpurun_out/f_tests.log
for v in base old lane t128x4 t128x4lane t192x3; do
  lib=build/variants/libellc_gn_$v.so
  [ $v = base ] && lib=egomotion_with_local_loop_closures_b200/libellc_gn.so
  ELLC_LIB=$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/f_var_$v.json 2> gpurun_out/f_var_$v.err
  echo "$v rc=$? $(python -c "import json;j=json.load(open('gpurun_out/f_var_$v.json'));print(round(j['value']), j['roofline']['kernel_ms_per_launch'])" 2>&1 | tail -1)"
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
head -c 2500 gpurun_out/f_bench.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/f_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_track -s 2 -c 1 -f -o gpurun_out/prof_r2_f python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/f_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
