#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "exchange" 2>&1 | tail -3
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/e_${N}_$name.json 2> gpurun_out/e_${N}_$name.err
  echo "$name rc=$? $(python -c "import json;j=json.load(open('gpurun_out/e_${N}_$name.json'));print(round(j['value']), round(j['ms_per_step'],3), 'e2e', j['e2e'] and round(j['e2e']['value']), j.get('per_rank'))" 2>&1 | tail -1)"
}
run p2p
run p2p_all --exchange-root -1
