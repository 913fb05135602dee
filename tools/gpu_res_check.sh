#!/bin/bash
# residual-join fusion: new parity tests + ResNet-50 bench A/B (fuse_eltwise=1 default vs 0)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "residual or resnet or reduce" > gpurun_out/res_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/res_pytest.log
for o in "" "fuse_eltwise=0"; do
for prec in fp32 bf16; do
  B200_FWD_OPTS=$o python bench.py --net resnet50 --batch 32 --prec $prec --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/res_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('resnet50 $prec [$o]', 'value %.0f'%d['value'],'ms %.4f'%d['ms_per_step'],'e2e %.0f'%d['e2e']['value'],'launches/step', d['gpu_launches']/d['steps'])" || tail -5 gpurun_out/res_bench.err
done; done
