#!/bin/bash
# round 2, call A: the whole GPU suite (incl. the new full-size config tests), the contraction-kernel diagnosis, the default bench line
mkdir -p gpurun_out
T=${1:-r02a}
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${T}_pytest.log
timeout 600 python tools/diag_conv.py > gpurun_out/${T}_diag.txt 2>&1; echo "diag rc=$?"; cat gpurun_out/${T}_diag.txt | cut -c1-400
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/${T}_bench.json
