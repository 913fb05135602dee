#!/bin/bash
# fused LRN + max-pool kernel: its tests, the AlexNet bench line with the per-call list, bf16 line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -x -q -k "lrn_inside or alexnet" 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline --no-other-configs > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/q_bench.json"))
print("value %.0f ms %.4f e2e %.0f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
for c in d["per_call"]: print("  %-24s call %.4f kernel %.4f" % (c["func"], c["call_ms"], c["kernel_ms"]))
PY
timeout 300 python bench.py --no-cpu-baseline --no-other-configs --prec bf16 2>/dev/null | cut -c1-200
