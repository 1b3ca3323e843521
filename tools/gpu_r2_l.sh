#!/bin/bash
# Round 2: L2 prefetch distance of the loop-closure record streams (A/B), depth-2 host pipeline of the bench (forward, with e2e).
mkdir -p gpurun_out
for v in base lcpf0 lcpf2 lcpf8 lcpf16; do
  lib=build/variants/libellc_gn_$v.so
  [ $v = base ] && lib=egomotion_with_local_loop_closures_b200/libellc_gn.so
  ELLC_LIB=$lib timeout 600 python bench.py --lc-mode const_weight --steps 10 --warmup 3 > gpurun_out/l_lc_$v.json 2> gpurun_out/l_lc_$v.err
  echo "lc $v rc=$? $(python -c "import json;j=json.load(open('gpurun_out/l_lc_$v.json'));print(round(j['value']), round(j['ms_per_step'],3), j['roofline']['kernel_ms_per_launch'], j.get('host_ms_per_step',{}).get('enqueue'))" 2>&1 | tail -1)"
done
for r in 1 2; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/l_fwd_$r.json 2> gpurun_out/l_fwd_$r.err
  echo "fwd $r rc=$? $(python -c "import json;j=json.load(open('gpurun_out/l_fwd_$r.json'));h=j.get('host_ms_per_step',{});print(round(j['value']), round(j['ms_per_step'],3), j['roofline']['kernel_ms_per_launch'], 'e2e', round(j['e2e']['value']), h.get('enqueue'), h.get('fetch'))" 2>&1 | tail -1)"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "loop_closure or lc or exchange" 2>&1 | tail -3
