#!/bin/bash
# ncu launch lists (graph nodes; duration, DRAM bytes, tensor-pipe activity per kernel) for "net batch prec skip count" tuples
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --graph-profiling node -s $4 -c $5 --csv --log-file gpurun_out/launches_$1_$3.csv python bench.py --net $1 --batch $2 --prec $3 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu launches $cfg rc=$? rows=$(wc -l < gpurun_out/launches_$1_$3.csv)"
done
