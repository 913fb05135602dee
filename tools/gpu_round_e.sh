#!/bin/bash
mkdir -p gpurun_out
T=${1:-r01i}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['achieved'])
for r in d['per_call']: print('  %-24s call %.4f kernel %.4f'%(r['func'],r['call_ms'],r['kernel_ms']))
PY
B200_FWD_OPTS="use_pdl=0" python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no-pdl value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --graph-profiling node -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
