for g in none peer_launch peer; do
B200_BENCH_GATHER=$g python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-other-configs 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$g value %.0f ms %.4f' % (d['value'], d['ms_per_step']))"
done
