#!/bin/bash
# Round 2: loop-closure pixel loop with cp.async-staged texel taps (A/B against the register version, occupancy variants), pipelined
# loop-closure bench mode, host-stall diagnostics of the default bench, parity tests.
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/k_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/k_tests.log
tail -12 gpurun_out/k_tests.log
for v in base lcsync lca4 lca2; do
  lib=build/variants/libellc_gn_$v.so
  [ $v = base ] && lib=egomotion_with_local_loop_closures_b200/libellc_gn.so
  ELLC_LIB=$lib timeout 600 python bench.py --lc-mode const_weight --steps 10 --warmup 3 > gpurun_out/k_lc_$v.json 2> gpurun_out/k_lc_$v.err
  echo "lc $v rc=$? $(python -c "import json;j=json.load(open('gpurun_out/k_lc_$v.json'));print(round(j['value']), round(j['ms_per_step'],3), j['roofline']['kernel_ms_per_launch'], j.get('host_ms_per_step',{}).get('enqueue'))" 2>&1 | tail -1)"
done
for r in 1 2 3; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/k_fwd_$r.json 2> gpurun_out/k_fwd_$r.err
  echo "fwd $r rc=$? $(python -c "import json;j=json.load(open('gpurun_out/k_fwd_$r.json'));h=j.get('host_ms_per_step',{});print(round(j['value']), round(j['ms_per_step'],3), j['roofline']['kernel_ms_per_launch'], h.get('enqueue'), h.get('fetch'))" 2>&1 | tail -1)"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_track_lc -s 2 -c 1 -f -o gpurun_out/prof_r2_lc6 python bench.py --lc-mode const_weight --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/k_ncu_lc.log 2>&1; echo "ncu lc rc=$?"
