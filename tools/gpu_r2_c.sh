#!/bin/bash
# Round 2, kernel variants (build/variants/*.so, selected with ELLC_LIB) on the default bench workload + the tests fixed after pass b.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_shim.py tests/test_gpu_parity.py -m gpu -q -k "surface or config1 or pair_list or exchange or lm_lambda or prepare_calls" > gpurun_out/c_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c_tests.log
tail -25 gpurun_out/c_tests.log
for v in base pf2 pf4 unz pf3unz; do
  lib=build/variants/libellc_gn_$v.so
  [ $v = base ] && lib=egomotion_with_local_loop_closures_b200/libellc_gn.so
  ELLC_LIB=$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c_var_$v.json 2> gpurun_out/c_var_$v.err
  echo "$v rc=$? $(python -c "import json;j=json.load(open('gpurun_out/c_var_$v.json'));print(round(j['value']), j['roofline']['kernel_ms_per_launch'])")"
done
