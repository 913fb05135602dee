#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02e}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -o gpurun_out/${T}_k1 python tools/diag_one.py k1 bf16 r1,sk4_im2col_dp > gpurun_out/${T}_k1.log 2>&1; echo "ncu k1 rc=$?"; tail -3 gpurun_out/${T}_k1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -o gpurun_out/${T}_conv5 python tools/diag_one.py conv5 bf16 r1,sk4_im2col_dp,sk4_halo_dp > gpurun_out/${T}_conv5.log 2>&1; echo "ncu conv5 rc=$?"; tail -4 gpurun_out/${T}_conv5.log
ls -la gpurun_out/${T}*
