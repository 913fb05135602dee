#!/bin/bash
mkdir -p gpurun_out
python tools/fc_experiments.py bf16 2>&1 | head -2
timeout 900 python -m pytest tests -m gpu -x -q -k "conv_edge or golden or sgemm or alexnet or tiny or concat" > gpurun_out/fc_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/fc_pytest.log
bash tools/gpu_percall.sh "alexnet_ng_conv 32 fp32" "googlenet_conv 64 bf16" "alexnet_ng_conv 32 bf16"
