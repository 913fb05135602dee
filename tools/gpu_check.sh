#!/bin/bash
mkdir -p gpurun_out
T=${1:-r01k}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
for cfg in "alexnet_ng_conv 32 fp32" "alexnet_ng_conv 32 bf16" "nin_imagenet 32 fp32" "googlenet_conv 64 bf16" "googlenet_conv 64 fp32" "resnet50 32 fp32" "resnet50 32 bf16"; do
  set -- $cfg
  python bench.py --net $1 --batch $2 --prec $3 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', 'value %.0f'%d['value'],'ms %.4f'%d['ms_per_step'],'e2e %.0f'%d['e2e']['value'],'roof %.1f'%d['roofline']['achieved'], 'launches/step', d['gpu_launches']/d['steps'])"
done
