#!/bin/bash
mkdir -p gpurun_out
T=${1:-r01g}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['achieved'])
for r in d['per_call']: print('  %-24s call %.4f kernel %.4f'%(r['func'],r['call_ms'],r['kernel_ms']))
PY
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --prec bf16 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bf16 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"
for prec in fp32 bf16; do echo "== c2 $prec"; python tools/ops_prof.py --ops-fn ops/c2-alexnet-ng-b32-convs.txt --prec $prec --no-check 2>&1 | cut -c1-130; done
echo "== c3 bf16"; python tools/ops_prof.py --ops-fn ops/c3-conv-ops-small.txt --prec bf16 --no-check 2>&1 | cut -c1-130
