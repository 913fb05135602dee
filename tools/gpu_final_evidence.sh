bash tools/gpu_final_bench.sh 2>&1 | tail -12
T=fin
cap() { local name=$1 re=$2 cnt=$3; shift 3
  timeout 900 ncu --set full --clock-control none -k regex:"$re" -c $cnt -o /tmp/${T}_${name} -f "$@" > gpurun_out/${T}_ncu_${name}.log 2>&1
  ncu -i /tmp/${T}_${name}.ncu-rep --page raw --csv > /tmp/${T}_${name}_raw.csv 2>/dev/null
  python tools/ncu_summarise.py /tmp/${T}_${name}_raw.csv "ncu --set full: ${name} ($*)" > gpurun_out/${T}_ncu_${name}.md
  if [ "$name" = "conv_fp32" ]; then cp /tmp/${T}_${name}_raw.csv gpurun_out/${T}_ncu_${name}_raw.csv; fi
}
cap conv_fp32 igemm 8 python tools/ops_prof.py --ops-fn ops/c2-alexnet-ng-b32-convs.txt --prec fp32 --iters 1 --warmup 0 --no-check
cap conv_bf16 igemm 8 python tools/ops_prof.py --ops-fn ops/c2-alexnet-ng-b32-convs.txt --prec bf16 --iters 1 --warmup 0 --no-check
cap pointwise 'lrn|pool|pack|absmax|finalize|splitk|reduce|l1max' 30 python bench.py --steps 1 --warmup 3 --no-cpu-baseline
ls gpurun_out | grep fin_ncu
