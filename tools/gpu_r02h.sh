#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02h}
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "sk4 or cluster_multicast or residual or 16bit" ) > gpurun_out/${T}_pytest_sk4.log 2>&1; echo "pytest sk4 rc=$?"; tail -6 gpurun_out/${T}_pytest_sk4.log | cut -c1-300
timeout 600 python tools/diag_conv.py > gpurun_out/${T}_diag.txt 2>&1; echo "diag rc=$?"; grep -v "role stamps" gpurun_out/${T}_diag.txt | cut -c1-420
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --no-other-configs > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${T}_bench.json
