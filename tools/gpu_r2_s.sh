#!/bin/bash
# Round 2: selection records prefetched into registers (ELLC_REC_REGS: 1 = L2-only loads, 2 = through L1) instead of the cp.async ring.
mkdir -p gpurun_out
for v in base recregs recregs2; do
  lib=build/variants/libellc_gn_$v.so
  [ $v = base ] && lib=egomotion_with_local_loop_closures_b200/libellc_gn.so
  ELLC_LIB=$lib timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/s_var_$v.json 2> gpurun_out/s_var_$v.err
  echo "$v rc=$? $(python -c "import json;j=json.load(open('gpurun_out/s_var_$v.json'));print(round(j['value']), round(j['ms_per_step'],3), j['roofline']['kernel_ms_per_launch'])" 2>&1 | tail -1)"
done
ELLC_LIB=build/variants/libellc_gn_recregs.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "track_end_to_end or reference_own or config1 or weights" 2>&1 | tail -3
