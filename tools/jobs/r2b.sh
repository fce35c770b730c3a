#!/bin/bash
# round-2 job B: validation of the folded / fused encoder path at the BASELINE shapes + bench A/B
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/r2b_tests.log
tail -6 gpurun_out/r2b_tests.log
(DFSFM_ENC_FUSED=1 timeout 600 python -m pytest tests/test_coarse_gpu.py tests/test_baseline_shapes_gpu.py tests/test_fullsize_gpu.py tests/test_stage_gpu.py -q 2>&1 | tail -30) > gpurun_out/r2b_tests_fused.log
tail -6 gpurun_out/r2b_tests_fused.log
(DFSFM_KV_EPI=0 timeout 600 python -m pytest tests/test_coarse_gpu.py tests/test_baseline_shapes_gpu.py -q 2>&1 | tail -30) > gpurun_out/r2b_tests_kvepi0.log
tail -4 gpurun_out/r2b_tests_kvepi0.log
for v in "0 1 0" "1 1 0" "1 1 1"; do set -- $v
  DFSFM_ATTN_FOLD=$1 DFSFM_KV_EPI=$2 DFSFM_ENC_FUSED=$3 timeout 300 python bench.py --steps 5 --warmup 3 --skip-hp2 --skip-post --skip-img --skip-cpu > gpurun_out/r2b_bench_fold$1_kvepi$2_fused$3.json 2> gpurun_out/r2b_bench_fold$1_kvepi$2_fused$3.err
done
