#!/bin/bash
# Round 2: the sporadic host stalls appear only at the full workload (512 frames / 32 keyframes): which half of the enqueue, with / without the
# NVML sampler thread, pipeline depth 1 / 2.
mkdir -p gpurun_out
nproc > gpurun_out/n_host.txt; free -m >> gpurun_out/n_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/n_host.txt 2>&1; cat /sys/fs/cgroup/memory.max >> gpurun_out/n_host.txt 2>&1; uptime >> gpurun_out/n_host.txt
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 60 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/n_$name.json 2> gpurun_out/n_$name.err
  echo "$name rc=$? $(python -c "import json;j=json.load(open('gpurun_out/n_$name.json'));h=j.get('host_ms_per_step',{});print(round(j['value']), round(j['ms_per_step'],3), round(j['roofline']['kernel_ms_per_launch'],3), {k:(round(v['median'],2),round(v['max'],1),v['argmax_step']) for k,v in h.items() if isinstance(v,dict)})" 2>&1 | tail -1)"
}
run d2_a ELLC_PIPE_DEPTH=2
run d2_noclk ELLC_PIPE_DEPTH=2 ELLC_NO_CLOCKS=1
run d1 ELLC_PIPE_DEPTH=1
run d2_b ELLC_PIPE_DEPTH=2
cat gpurun_out/n_host.txt
