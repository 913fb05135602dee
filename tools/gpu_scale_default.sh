#!/bin/bash
# the driver's scaling command at N GPUs (default flags: main line + other_configs). usage: gpu_scale_default.sh N
N=$1
mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 2>gpurun_out/scale_default_${N}.err | tail -1 > gpurun_out/scale_default_${N}.json
python - <<PY
import json
d=json.load(open('gpurun_out/scale_default_${N}.json'))
print('N=$N value %.0f ms %.4f e2e %.0f %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('parallelism')))
for k, o in (d.get("other_configs") or {}).items(): print("   other", k, "value %.0f" % o["value"], "e2e %.0f" % o["e2e"]["value"])
PY
tail -3 gpurun_out/scale_default_${N}.err
