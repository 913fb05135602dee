#!/bin/bash
# end-of-round evidence: full GPU test suite, default bench (with cpu_baseline) + reference arm, per-op sweeps with oracle check, ncu launch list + --set full summaries (exported on the box)
mkdir -p gpurun_out
T=${1:-r01z}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${T}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/${T}_bench_reference.json
python bench.py --prec bf16 --no-cpu-baseline > gpurun_out/${T}_bench_bf16.json 2>/dev/null
python bench.py --net googlenet_conv --batch 64 --prec bf16 --steps 30 --no-cpu-baseline > gpurun_out/${T}_bench_googlenet_bf16.json 2>/dev/null
python bench.py --net resnet50 --batch 32 --steps 30 --no-cpu-baseline > gpurun_out/${T}_bench_resnet50_fp32.json 2>/dev/null
python bench.py --net nin_imagenet --steps 30 --no-cpu-baseline > gpurun_out/${T}_bench_nin_fp32.json 2>/dev/null
for cfg in "c1-sgemm-ops-tiny fp32" "c2-alexnet-ng-b32-convs fp32" "c2-alexnet-ng-b32-convs bf16" "c3-conv-ops-small fp32" "c3-conv-ops-small fp16" "c3-conv-ops-small bf16"; do
  set -- $cfg
  short=$(echo $1 | cut -d- -f1)
  python tools/ops_prof.py --ops-fn ops/$1.txt --prec $2 --out gpurun_out/${T}_ops_prof_${short}_$2.md --wisdom-out gpurun_out/${T}_wisdom_${short}_$2.wis > gpurun_out/${T}_ops_prof_${short}_$2.log 2>&1; echo "ops_prof $cfg rc=$?"
done
rm -f gpurun_out/${T}_wisdom_c1_fp32.wis   # 8192^2 digests are fine but the file is not needed
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --graph-profiling node -s 300 -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
cap() { local name=$1 re=$2 cnt=$3; shift 3
  timeout 900 ncu --set full --clock-control none -k regex:"$re" -c $cnt -o /tmp/${T}_${name} -f "$@" > gpurun_out/${T}_ncu_${name}.log 2>&1
  ncu -i /tmp/${T}_${name}.ncu-rep --page raw --csv > /tmp/${T}_${name}_raw.csv 2>/dev/null
  python tools/ncu_summarise.py /tmp/${T}_${name}_raw.csv "ncu --set full: ${name} ($*)" > gpurun_out/${T}_ncu_${name}.md
  if [ "$name" = "conv_fp32" ]; then cp /tmp/${T}_${name}_raw.csv gpurun_out/${T}_ncu_${name}_raw.csv; fi
}
cap conv_fp32 igemm 8 python tools/ops_prof.py --ops-fn ops/c2-alexnet-ng-b32-convs.txt --prec fp32 --iters 1 --warmup 0 --no-check
cap conv_bf16 igemm 8 python tools/ops_prof.py --ops-fn ops/c2-alexnet-ng-b32-convs.txt --prec bf16 --iters 1 --warmup 0 --no-check
cap c3_fp16 igemm 10 python tools/ops_prof.py --ops-fn ops/c3-conv-ops-small.txt --prec fp16 --iters 1 --warmup 0 --no-check
cap pointwise 'lrn|pool|pack|absmax|finalize|splitk' 24 python bench.py --steps 1 --warmup 3 --no-cpu-baseline
du -sh gpurun_out; ls gpurun_out | grep ${T} | head -50
