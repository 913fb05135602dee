#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02d}
timeout 600 python tools/diag_conv.py > gpurun_out/${T}_diag.txt 2>&1; echo "diag rc=$?"; grep -v "role stamps" gpurun_out/${T}_diag.txt | cut -c1-420
