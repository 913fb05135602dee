#!/bin/bash
# round-2 evidence, part A: the driver's round-end sequence (GPU tests, smoke, default bench line, reference arm) + the other nets / precisions
mkdir -p gpurun_out
T=r02fin
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${T}_pytest.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference rc=$?"; cut -c1-400 gpurun_out/${T}_bench_reference.json
for cfg in "alexnet_ng_conv 32 bf16" "alexnet_ng_conv 32 fp16" "nin_imagenet 32 fp32" "nin_imagenet 32 bf16" "googlenet_conv 64 fp32" "resnet50 32 bf16"; do
  set -- $cfg
  timeout 300 python bench.py --net $1 --batch $2 --prec $3 --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/${T}_bench_$1_$3.json 2>/dev/null
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_bench_*.json')):
    try: d=json.load(open(f))
    except Exception as e: print(f, 'unreadable', e); continue
    if 'value' not in d: print(f, d); continue
    print('%-44s value %9.0f ms %.4f e2e %9.0f frac %s launches/step %s' % (f.split('${T}_bench_')[1][:-5], d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('roofline') or {}).get('frac'), d.get('gpu_launches', 0) / max(1, d['steps'])))
    for k, o in (d.get('other_configs') or {}).items(): print('    %-40s value %9.0f ms %.4f e2e %9.0f' % (k, o['value'], o['ms_per_step'], o['e2e']['value']))
    if d.get('cpu_baseline'): print('    cpu_baseline', json.dumps(d['cpu_baseline'])[:300])
    if d.get('sustained'): print('    sustained', d['sustained'])
PY
