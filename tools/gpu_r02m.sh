#!/bin/bash
# 2 GPUs: the peer-memory logits gather + C-ABI NCCL broadcast through bench.py, and the NCCL all-gather A/B
mkdir -p gpurun_out
T=${1:-r02m}
N=${2:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --no-other-configs > gpurun_out/${T}_bench_n${N}.json 2> gpurun_out/${T}_bench_n${N}.err; echo "bench N=$N rc=$?"; cut -c1-700 gpurun_out/${T}_bench_n${N}.json; tail -5 gpurun_out/${T}_bench_n${N}.err | cut -c1-300
B200_BENCH_GATHER=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 --no-other-configs > gpurun_out/${T}_bench_n${N}_nccl.json 2> gpurun_out/${T}_bench_n${N}_nccl.err; echo "bench nccl N=$N rc=$?"; cut -c1-400 gpurun_out/${T}_bench_n${N}_nccl.json; tail -3 gpurun_out/${T}_bench_n${N}_nccl.err | cut -c1-300

