#!/bin/bash
# bench lines (with the per-call eager profile they carry) for the nets named on the command line: "net batch prec" triples
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  python bench.py --net $1 --batch $2 --prec $3 --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/percall_$1_$3.err | tail -1 > gpurun_out/percall_$1_$3.json
  python -c "import json; d=json.load(open('gpurun_out/percall_$1_$3.json')); print('$cfg', 'value %.0f'%d['value'],'ms %.4f'%d['ms_per_step'],'e2e %.0f'%d['e2e']['value'])" || tail -5 gpurun_out/percall_$1_$3.err
done
