#!/bin/bash
# Round 2: where do the sporadic host stalls of the bench loop come from?  Small workload, many steps, NVML polling period 5 / 50 ms / off.
mkdir -p gpurun_out
for rep in 1 2 3; do
for per in 5 50 1000000; do
  ELLC_CLOCK_PERIOD_MS=$per timeout 300 python bench.py --frames 128 --keyframes 8 --steps 300 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/m_stall_${per}_$rep.json 2> gpurun_out/m_stall_${per}_$rep.err
  echo "period $per rep $rep rc=$? $(python -c "import json;j=json.load(open('gpurun_out/m_stall_${per}_$rep.json'));h=j.get('host_ms_per_step',{});print(round(j['value']), round(j['ms_per_step'],3), j['roofline']['kernel_ms_per_launch'], h.get('enqueue'), h.get('fetch'), j['clocks']['samples'])" 2>&1 | tail -1)"
done
done
for v in base recef; do
  lib=build/variants/libellc_gn_$v.so
  [ $v = base ] && lib=egomotion_with_local_loop_closures_b200/libellc_gn.so
  ELLC_LIB=$lib timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/m_var_$v.json 2> gpurun_out/m_var_$v.err
  echo "$v rc=$? $(python -c "import json;j=json.load(open('gpurun_out/m_var_$v.json'));h=j.get('host_ms_per_step',{});print(round(j['value']), round(j['ms_per_step'],3), j['roofline']['kernel_ms_per_launch'], h.get('enqueue'))" 2>&1 | tail -1)"
done
