#!/bin/bash
# Round 2, two GPUs: the in-kernel result exchange over peer memory (CUDA IPC) against the NCCL gather, gather vs all-gather.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/d_topo.txt 2>&1
run() { # name, extra args
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 "$@" > gpurun_out/d_$name.json 2> gpurun_out/d_$name.err
  echo "$name rc=$? $(python -c "import json;j=json.load(open('gpurun_out/d_$name.json'));print(round(j['value']), round(j['ms_per_step'],3), 'k_ms', round(j['roofline']['kernel_ms_per_launch'],3), 'share', round(j['roofline']['kernel_share_of_step'],3), 'e2e', j['e2e'] and round(j['e2e']['value']), j['config']['parallelism'][-90:])" 2>&1 | tail -1)"
  grep -i "fall\|error\|unavailable" gpurun_out/d_$name.err | head -5
}
run p2p_root0
run p2p_all --exchange-root -1
run nccl --exchange nccl
run p2p_lc --lc-mode const_weight
