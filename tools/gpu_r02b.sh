#!/bin/bash
# round 2, call B: the round-2 kernel (igemm4.cuh) -- its own tests first, then the whole GPU suite, the diagnosis runs and the bench
mkdir -p gpurun_out
T=${1:-r02b}
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "sk4 or cluster_multicast or tap_reuse" ) > gpurun_out/${T}_pytest_sk4.log 2>&1; echo "pytest sk4 rc=$?"; tail -15 gpurun_out/${T}_pytest_sk4.log | cut -c1-400
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${T}_pytest.log | cut -c1-400
timeout 600 python tools/diag_conv.py > gpurun_out/${T}_diag.txt 2>&1; echo "diag rc=$?"; grep -v "role stamps" gpurun_out/${T}_diag.txt | cut -c1-420
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${T}_bench.json
timeout 300 python bench.py --no-cpu-baseline --prec bf16 > gpurun_out/${T}_bench_bf16.json 2> gpurun_out/${T}_bench_bf16.err; echo "bench bf16 rc=$?"; cut -c1-300 gpurun_out/${T}_bench_bf16.json
