#!/bin/bash
# what the driver runs at round end: GPU test suite, smoke(), default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/end_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/end_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/end_bench.json 2>gpurun_out/end_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/end_bench.json
