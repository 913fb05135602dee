#!/bin/bash
# gpurun call: parity tests for the tap-reuse kernel, per-op sweeps (taps on/off), ncu captures exported to CSV on the box
mkdir -p gpurun_out
T=r01e
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -15 gpurun_out/${T}_pytest.log
for prec in fp32 bf16; do
  for o in "use_taps=0" "use_taps=1" "use_taps=1,taps_2cta=0" "use_taps=1,taps_2cta=1"; do
    echo "== c2 $prec $o"; python tools/ops_prof.py --ops-fn ops/c2-alexnet-ng-b32-convs.txt --prec $prec --no-check --opts "$o" 2>&1 | grep -E "conv (5x5|3x3)" | cut -c1-150
  done
done
echo "== c3 bf16 chunk sweep (im2col + 1x1 paths too)"
for o in "acc_chunk_kblks_16=4" "acc_chunk_kblks_16=32"; do echo "-- $o"; python tools/ops_prof.py --ops-fn ops/c3-conv-ops-small.txt --prec bf16 --no-check --opts "$o" 2>&1 | cut -c1-150; done
# ncu: conv3 + conv4 of C2 (lines 3,4), fp32 and bf16, taps kernel; export raw pages here, keep only the CSVs
sed -n 3,4p ops/c2-alexnet-ng-b32-convs.txt > /tmp/c2_34.txt
for prec in fp32 bf16; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -c 2 -o /tmp/${T}_taps_${prec} -f python tools/ops_prof.py --ops-fn /tmp/c2_34.txt --prec $prec --iters 1 --warmup 0 --no-check > gpurun_out/${T}_ncu_${prec}.log 2>&1
  ncu -i /tmp/${T}_taps_${prec}.ncu-rep --page raw --csv > gpurun_out/${T}_taps_${prec}_raw.csv 2>/dev/null
  ncu -i /tmp/${T}_taps_${prec}.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/${T}_taps_${prec}_source.csv 2>/dev/null
  ls -la /tmp/${T}_taps_${prec}.ncu-rep
done
cp /tmp/${T}_taps_bf16.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out
