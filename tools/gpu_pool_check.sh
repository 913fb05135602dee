#!/bin/bash
# pool kernel rework: parity tests that touch pooling + bench lines of the nets it matters for
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "pool or tiny or googlenet_conv_b2_all_nodes or alexnet" > gpurun_out/pool_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pool_pytest.log
bash tools/gpu_percall.sh "alexnet_ng_conv 32 fp32" "googlenet_conv 64 bf16" "resnet50 32 fp32" "nin_imagenet 32 fp32"
