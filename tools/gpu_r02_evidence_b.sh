#!/bin/bash
# round-2 evidence, part B: per-op sweeps with the vendor comparator, ncu launch list of the default bench command, ncu --set full of the kernel families
mkdir -p gpurun_out
T=r02fin
for cfg in "c1-sgemm-ops-tiny fp32" "c2-alexnet-ng-b32-convs fp32" "c2-alexnet-ng-b32-convs bf16" "c3-conv-ops-small fp32" "c3-conv-ops-small fp16" "c3-conv-ops-small bf16"; do
  set -- $cfg
  timeout 600 python tools/ops_prof.py --ops-fn ops/$1.txt --prec $2 --compare --out gpurun_out/${T}_ops_prof_${1%%-*}_$2.json > gpurun_out/${T}_ops_prof_${1%%-*}_$2.md 2>gpurun_out/${T}_ops_prof_${1%%-*}_$2.err; echo "ops_prof $cfg rc=$?"; tail -4 gpurun_out/${T}_ops_prof_${1%%-*}_$2.md | cut -c1-220
done
# launch list of the SAME command the bench line comes from (graph nodes)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --graph-profiling node -s 60 -c 40 --csv --log-file gpurun_out/${T}_launches_alexnet_ng_conv_fp32.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launchlist_md.py gpurun_out/${T}_launches_alexnet_ng_conv_fp32.csv "AlexNet-ng B=32 fp32-parity: python bench.py --steps 3 --warmup 3 (graph nodes), round 2 final" > gpurun_out/${T}_launches_alexnet_ng_conv_fp32.md; tail -3 gpurun_out/${T}_launches_alexnet_ng_conv_fp32.md | cut -c1-300
cap() { local name=$1 re=$2 cnt=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$re" -c $cnt -o /tmp/${T}_${name} -f "$@" > gpurun_out/${T}_ncu_${name}.log 2>&1
  ncu -i /tmp/${T}_${name}.ncu-rep --page raw --csv > /tmp/${T}_${name}_raw.csv 2>/dev/null
  python tools/ncu_summarise.py /tmp/${T}_${name}_raw.csv "ncu --set full: ${name} ($*)" > gpurun_out/${T}_ncu_${name}.md; echo "cap $name rows=$(wc -l < gpurun_out/${T}_ncu_${name}.md)"
  if [ "$name" = "step_fp32" ]; then cp /tmp/${T}_${name}_raw.csv gpurun_out/${T}_ncu_${name}_raw.csv; fi
}
# one whole forward of the bench command, every kernel (10 per step): skip the warm-up forwards
cap step_fp32 'igemm|fc_chain|lrn_maxpool|pool|absmax_pack' 20 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs
cap step_bf16 'igemm|fc_chain|lrn_maxpool|pool|pack' 20 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs --prec bf16
ls -la gpurun_out | grep ${T}_ | awk '{print $5, $9}' | tail -30
