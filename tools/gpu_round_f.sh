#!/bin/bash
# ncu --set full captures of every kernel family on the hot path, exported to CSV/markdown ON the box (the .ncu-rep files are too large to bring back)
mkdir -p gpurun_out
T=${1:-r01j}
cap() { # name, kernel regex, count, command...
  local name=$1 re=$2 cnt=$3; shift 3
  timeout 900 ncu --set full --clock-control none -k regex:"$re" -c $cnt -o /tmp/${T}_${name} -f "$@" > gpurun_out/${T}_ncu_${name}.log 2>&1
  ncu -i /tmp/${T}_${name}.ncu-rep --page raw --csv > /tmp/${T}_${name}_raw.csv 2>/dev/null
  python tools/ncu_summarise.py /tmp/${T}_${name}_raw.csv "ncu --set full: ${name} ($*)" > gpurun_out/${T}_ncu_${name}.md
  # keep the full raw page for the dominant kernel only
  if [ "$name" = "conv_fp32" ]; then cp /tmp/${T}_${name}_raw.csv gpurun_out/${T}_ncu_${name}_raw.csv; fi
  tail -n +4 gpurun_out/${T}_ncu_${name}.md | cut -c1-260
}
cap conv_fp32 igemm 8 python tools/ops_prof.py --ops-fn ops/c2-alexnet-ng-b32-convs.txt --prec fp32 --iters 1 --warmup 0 --no-check
cap conv_bf16 igemm 8 python tools/ops_prof.py --ops-fn ops/c2-alexnet-ng-b32-convs.txt --prec bf16 --iters 1 --warmup 0 --no-check
cap c3_fp16 igemm 10 python tools/ops_prof.py --ops-fn ops/c3-conv-ops-small.txt --prec fp16 --iters 1 --warmup 0 --no-check
cap pointwise 'lrn|pool|pack|absmax|finalize|splitk' 24 python bench.py --steps 1 --warmup 3 --no-cpu-baseline
du -sh gpurun_out
echo "== googlenet bf16 B=64"
python bench.py --net googlenet_conv --batch 64 --prec bf16 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-700
echo "== nin fp32 B=32"
python bench.py --net nin_imagenet --batch 32 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-400
