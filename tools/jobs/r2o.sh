#!/bin/bash
# round-2 job O (1 GPU): tcgen05.cp experiment (A operand smem -> TMEM with row-shifted SW128 descriptors) + the tile-stealing subprocess test
mkdir -p gpurun_out
(nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I detectorfreesfm_b200/csrc tools/exp_tmem_cp.cu -o /tmp/exp_tmem_cp 2>/dev/null && timeout 60 /tmp/exp_tmem_cp) > gpurun_out/r2o_tmem_cp.log 2>&1
echo "rc=$?" >> gpurun_out/r2o_tmem_cp.log
cat gpurun_out/r2o_tmem_cp.log
(timeout 900 python -m pytest tests/test_coarse_gpu.py -m gpu -q -x -k "stealing or schedules" 2>&1 | tail -5) > gpurun_out/r2o_tests.log
tail -3 gpurun_out/r2o_tests.log
