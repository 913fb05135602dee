mkdir -p gpurun_out
python bench.py > gpurun_out/fin_alexnet_fp32.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin_reference.json 2>/dev/null
for cfg in "alexnet_ng_conv 32 bf16" "alexnet_ng_conv 32 fp16" "nin_imagenet 32 fp32" "nin_imagenet 32 bf16" "googlenet_conv 64 bf16" "googlenet_conv 64 fp32" "resnet50 32 fp32" "resnet50 32 bf16"; do
  set -- $cfg
  python bench.py --net $1 --batch $2 --prec $3 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/fin_$1_$3.json 2>/dev/null
done
for f in gpurun_out/fin_*.json; do python -c "
import json,sys
d=json.load(open('$f'))
print('%-40s value %9.0f  ms %.4f  e2e %9.0f  roof %.1f TF/s frac %.3f'%('$f'.split('fin_')[1][:-5], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('roofline',{}).get('achieved',0), d.get('roofline',{}).get('frac',0) or 0))"; done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --graph-profiling node -s 300 -c 300 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu launches rc=$?"
