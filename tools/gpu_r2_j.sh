#!/bin/bash
# Round 2, N GPUs of one box: in-kernel result exchange at N ranks (forward and loop-closure mode).  usage: gpu_r2_j.sh N
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi topo -m > gpurun_out/j_topo_$N.txt 2>&1
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/j_${N}_$name.json 2> gpurun_out/j_${N}_$name.err
  echo "$name rc=$? $(python -c "import json;j=json.load(open('gpurun_out/j_${N}_$name.json'));print(round(j['value']), round(j['ms_per_step'],3), 'k_ms', round(j['roofline']['kernel_ms_per_launch'],3), 'share', round(j['roofline']['kernel_share_of_step'],3), 'e2e', j['e2e'] and round(j['e2e']['value']), j.get('host_ms_per_step'))" 2>&1 | tail -1)"
  grep -i "fall\|error\|unavailable" gpurun_out/j_${N}_$name.err | head -5
}
run p2p --steps 20 --warmup 5
run p2p_lc --lc-mode const_weight --steps 10 --warmup 3 --no-cpu-baseline
