#!/bin/bash
# Round 2: loop-closure pixel loop software-pipelined in registers (ELLC_LC_PIPE) at 80 / 128 registers against the committed loop; parity of the variant.
mkdir -p gpurun_out
for v in base lcp3 lcp2; do
  lib=build/variants/libellc_gn_$v.so
  [ $v = base ] && lib=egomotion_with_local_loop_closures_b200/libellc_gn.so
  ELLC_LIB=$lib timeout 600 python bench.py --lc-mode const_weight --steps 20 --warmup 3 > gpurun_out/o_lc_$v.json 2> gpurun_out/o_lc_$v.err
  echo "lc $v rc=$? $(python -c "import json;j=json.load(open('gpurun_out/o_lc_$v.json'));print(round(j['value']), round(j['ms_per_step'],3), j['roofline']['kernel_ms_per_launch'], j.get('host_ms_per_step',{}).get('launch'))" 2>&1 | tail -1)"
done
ELLC_LIB=build/variants/libellc_gn_lcp3.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_host_shim.py -m gpu -q -k "loop_closure or lc or const_weight or shim" 2>&1 | tail -4
