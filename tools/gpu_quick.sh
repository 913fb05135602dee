#!/bin/bash
# quick regression: the whole GPU suite + the default bench line (no CPU legs) + bf16
mkdir -p gpurun_out
T=${1:-q}
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --no-other-configs > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('value %.0f ms %.4f e2e %.0f frac %.3f launches/step %.1f sustained %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches']/d['steps'], (d.get('sustained') or {}).get('value')))
for c in d['per_call']: print('  %-24s call %.4f kernel %.4f' % (c['func'], c['call_ms'], c['kernel_ms']))
PY
timeout 300 python bench.py --no-cpu-baseline --no-other-configs --prec bf16 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bf16 value %.0f ms %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
