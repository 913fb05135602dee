#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02l}
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${T}_pytest.log | cut -c1-300
for prec in fp32 bf16; do
  for o in "" "use_sk4=1"; do
    B200_FWD_OPTS=$o timeout 300 python bench.py --no-cpu-baseline --no-other-configs --prec $prec --steps 30 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('alexnet $prec [$o]', 'value %.0f'%d['value'],'ms %.4f'%d['ms_per_step'],'e2e %.0f'%d['e2e']['value'],'roof %.1f'%d['roofline']['achieved'], 'launches/step', d['gpu_launches']/d['steps'])"
  done
done
for o in "" "use_sk4=1"; do
  B200_FWD_OPTS=$o timeout 300 python bench.py --no-cpu-baseline --no-other-configs --net googlenet_conv --batch 64 --prec bf16 --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('googlenet bf16 [$o]', 'value %.0f'%d['value'],'ms %.4f'%d['ms_per_step'],'e2e %.0f'%d['e2e']['value'],'roof %.1f'%d['roofline']['achieved'], 'launches/step', d['gpu_launches']/d['steps'])"
  B200_FWD_OPTS=$o timeout 300 python bench.py --no-cpu-baseline --no-other-configs --net resnet50 --batch 32 --prec bf16 --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('resnet50 bf16 [$o]', 'value %.0f'%d['value'],'ms %.4f'%d['ms_per_step'],'e2e %.0f'%d['e2e']['value'],'roof %.1f'%d['roofline']['achieved'], 'launches/step', d['gpu_launches']/d['steps'])"
done
