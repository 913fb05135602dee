#!/bin/bash
# round-2 evidence, part C: the ncu half of part B alone (launch list of the bench command + ncu --set full of one forward), for a refresh at HEAD
mkdir -p gpurun_out
T=r02fin
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --graph-profiling node -s 60 -c 40 --csv --log-file gpurun_out/${T}_launches_alexnet_ng_conv_fp32.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > /dev/null 2>&1; echo "ncu launches rc=$?"
python tools/launchlist_md.py gpurun_out/${T}_launches_alexnet_ng_conv_fp32.csv "AlexNet-ng B=32 fp32-parity: python bench.py --steps 3 --warmup 3 (graph nodes), round 2 final" > gpurun_out/${T}_launches_alexnet_ng_conv_fp32.md; tail -3 gpurun_out/${T}_launches_alexnet_ng_conv_fp32.md | cut -c1-300
cap() { local name=$1 re=$2 cnt=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$re" -c $cnt -o /tmp/${T}_${name} -f "$@" > gpurun_out/${T}_ncu_${name}.log 2>&1
  ncu -i /tmp/${T}_${name}.ncu-rep --page raw --csv > /tmp/${T}_${name}_raw.csv 2>/dev/null
  python tools/ncu_summarise.py /tmp/${T}_${name}_raw.csv "ncu --set full: ${name} ($*)" > gpurun_out/${T}_ncu_${name}.md; echo "cap $name rows=$(wc -l < gpurun_out/${T}_ncu_${name}.md)"
  if [ "$name" = "step_fp32" ]; then cp /tmp/${T}_${name}_raw.csv gpurun_out/${T}_ncu_${name}_raw.csv; fi
}
cap step_fp32 'igemm|fc_chain|lrn_maxpool|pool|absmax_pack' 20 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs
cap step_bf16 'igemm|fc_chain|lrn_maxpool|pool|pack' 20 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs --prec bf16
