#!/bin/bash
# one gpurun call: GPU parity tests, default bench, ncu launch list, ncu --set full captures
mkdir -p gpurun_out
T=r01d
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
python bench.py --steps 50 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm -c 8 -o gpurun_out/${T}_igemm_c2_fp32 -f python tools/ops_prof.py --ops-fn ops/c2-alexnet-ng-b32-convs.txt --prec fp32 --iters 1 --warmup 0 --no-check > gpurun_out/${T}_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm -c 10 -o gpurun_out/${T}_igemm_c3_bf16 -f python tools/ops_prof.py --ops-fn ops/c3-conv-ops-small.txt --prec bf16 --iters 1 --warmup 0 --no-check > gpurun_out/${T}_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lrn|pool|pack|absmax|splitk' -c 24 -o gpurun_out/${T}_pointwise -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_pw.log 2>&1; echo "ncu pw rc=$?"
ls -la gpurun_out
