python bench.py --net resnet50 --batch 32 --prec fp32 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
rows=sorted(d['per_call'], key=lambda r:-r['call_ms'])
tot=sum(r['call_ms'] for r in d['per_call']); kt=sum(r['kernel_ms'] for r in d['per_call'])
print('resnet50 fp32 ms/step',d['ms_per_step'],'sum call',tot,'sum kernel',kt)
for r in rows[:14]: print('  %-34s call %.4f kernel %.4f gflop %.2f'%(r['func'],r['call_ms'],r['kernel_ms'],r['gflop']))
"
python bench.py --net googlenet_conv --batch 64 --prec bf16 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys,collections
d=json.loads(sys.stdin.read())
rows=sorted(d['per_call'], key=lambda r:-r['call_ms'])
tot=sum(r['call_ms'] for r in d['per_call']); kt=sum(r['kernel_ms'] for r in d['per_call'])
print('googlenet bf16 ms/step',d['ms_per_step'],'sum call',tot,'sum kernel',kt)
bykind=collections.Counter()
for r in d['per_call']: bykind[r['func'].split('__')[0]]+=r['call_ms']
print(dict(bykind))
for r in rows[:10]: print('  %-34s call %.4f kernel %.4f gflop %.2f'%(r['func'],r['call_ms'],r['kernel_ms'],r['gflop']))
"
