#!/bin/bash
# Run on the GPU box (through gpurun): launch list + full ncu capture of the dominant kernels.  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 1 200 > gpurun_out/ncu_launches.log 2>&1
# backbone conv GEMMs of the second image (launches 14..27 of gemm_tc2_kernel): layer1 convs first
timeout 300 ncu --set full --clock-control none --import-source on -k gemm_tc2_kernel -s 14 -c 5 -o gpurun_out/prof_conv python tools/profile_step.py 1 0 > gpurun_out/ncu_conv.log 2>&1
tail -1 gpurun_out/ncu_launches.log gpurun_out/ncu_conv.log
