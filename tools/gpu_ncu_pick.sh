#!/bin/bash
# one `ncu --set full` capture of selected kernels of a bench run, raw CSV exported on the box (reports are too big to travel)
# usage: gpu_ncu_pick.sh name 'kernel-regex' skip count -- bench args
name=$1; re=$2; skip=$3; cnt=$4; shift 5
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --graph-profiling node -k regex:"$re" -s $skip -c $cnt -o /tmp/pick_$name -f python bench.py "$@" --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/pick_${name}.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/pick_$name.ncu-rep --page raw --csv > gpurun_out/pick_${name}_raw.csv 2>/dev/null
python tools/ncu_summarise.py gpurun_out/pick_${name}_raw.csv "ncu --set full: $name" | cut -c1-260
ls -la gpurun_out/pick_${name}_raw.csv
