#!/bin/bash
# Round 2, one GPU: gather-early variant A/B, BASELINE config 5 (pair sweep 252 .. 65,536 pairs), configs 2 / 4, ncu capture of the loop-closure kernel.
mkdir -p gpurun_out
for v in base gearly; do
  lib=build/variants/libellc_gn_$v.so
  [ $v = base ] && lib=egomotion_with_local_loop_closures_b200/libellc_gn.so
  ELLC_LIB=$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/i_var_$v.json 2> gpurun_out/i_var_$v.err
  echo "$v rc=$? $(python -c "import json;j=json.load(open('gpurun_out/i_var_$v.json'));print(round(j['value']), j['roofline']['kernel_ms_per_launch'])" 2>&1 | tail -1)"
done
: > gpurun_out/i_sweep.jsonl
for fk in "28 8 9" "114 16 9" "512 32 9" "1024 32 9" "2048 64 9" "4096 64 9" "4096 64 16"; do
  set -- $fk
  timeout 1200 python bench.py --frames $1 --keyframes $2 --pairs-per-frame $3 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tee -a gpurun_out/i_sweep.jsonl | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; c=d['config']
print('frames %5d keyframes %3d pairs %6d : %8.0f tracks/s  step %7.2f ms  kernel %7.2f ms  frac %.3f' % (c['frames_per_gpu'], c['keyframes_per_gpu'], c['pairs_per_gpu_per_step'], d['value'], d['ms_per_step'], r['kernel_ms_per_launch'], r['frac']))"
done
timeout 600 python bench.py --config 720p_single --steps 50 --warmup 5 > gpurun_out/i_bench_720p.json 2> gpurun_out/i_bench_720p.err; echo "720p rc=$?"
timeout 600 python bench.py --config 1080p_stress --steps 10 --warmup 3 > gpurun_out/i_bench_1080p.json 2> gpurun_out/i_bench_1080p.err; echo "1080p rc=$?"
timeout 600 python bench.py --lc-mode const_weight --steps 10 --warmup 3 > gpurun_out/i_bench_lc.json 2> gpurun_out/i_bench_lc.err; echo "lc rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_track_lc -s 2 -c 1 -f -o gpurun_out/prof_r2_lc5 python bench.py --lc-mode const_weight --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/i_ncu_lc.log 2>&1; echo "ncu lc rc=$?"
