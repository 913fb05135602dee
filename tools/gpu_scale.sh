#!/bin/bash
# multi-GPU bench on one box: torchrun, one rank per GPU (NCCL). usage: gpu_scale.sh N
N=$1
mkdir -p gpurun_out
run() { # net batch prec
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --net $1 --batch $2 --prec $3 2>gpurun_out/scale_${N}_$1_$3.err | tail -1 > gpurun_out/scale_${N}_$1_$3.json
  python -c "import json; d=json.load(open('gpurun_out/scale_${N}_$1_$3.json')); print('N=$N', '$1', '$3', 'value %.0f'%d['value'], 'ms %.4f'%d['ms_per_step'], 'e2e %.0f'%d['e2e']['value'], d['config']['parallelism'])" || tail -5 gpurun_out/scale_${N}_$1_$3.err
}
nvidia-smi -L | head -8
run alexnet_ng_conv 32 fp32
if [ "$N" = "4" ] || [ "$N" = "8" ]; then run googlenet_conv 64 bf16; fi
if [ "$N" = "8" ]; then run resnet50 32 fp32; fi
