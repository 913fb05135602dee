#!/bin/bash
# A/B of the multi-GPU step loop on one box: logits gather one step late beside the next forward (default) vs serial (B200_BENCH_SERIAL_GATHER=1).
# usage: gpu_scale_ab.sh N
N=$1
mkdir -p gpurun_out
run() { # tag env
  env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/scale_ab_${N}_$1.err | tail -1 > gpurun_out/scale_ab_${N}_$1.json
  python -c "import json; d=json.load(open('gpurun_out/scale_ab_${N}_$1.json')); print('N=$N', '$1', 'value %.0f'%d['value'], 'ms %.4f'%d['ms_per_step'], 'e2e %.0f'%d['e2e']['value'])" || tail -15 gpurun_out/scale_ab_${N}_$1.err
}
run pipelined B200_BENCH_SERIAL_GATHER=0
run serial B200_BENCH_SERIAL_GATHER=1
