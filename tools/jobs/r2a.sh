#!/bin/bash
# round-2 job A: TMEM-A experiment, GPU test suite, first run of the fused encoder kernel (opt-in), bench A/B, ncu of HBM-class kernels
mkdir -p gpurun_out
timeout 60 ./tools/exp_tmem_a > gpurun_out/r2a_exp_tmem_a.log 2>&1; echo "exp_tmem_a rc=$?" >> gpurun_out/r2a_exp_tmem_a.log
cat gpurun_out/r2a_exp_tmem_a.log
(timeout 300 python -m pytest tests/test_coarse_gpu.py -x -q -k "transformer or end_to_end_pair" 2>&1 | tail -8) > gpurun_out/r2a_tests_quick.log
cat gpurun_out/r2a_tests_quick.log
(DFSFM_ENC_FUSED=1 timeout 180 python -m pytest tests/test_coarse_gpu.py -x -q -k "transformer or end_to_end_pair" 2>&1 | tail -15) > gpurun_out/r2a_tests_fused.log; echo "fused rc=$?" >> gpurun_out/r2a_tests_fused.log
cat gpurun_out/r2a_tests_fused.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2a_tests.log
for v in "0 1 0" "1 0 0" "1 1 0" "1 1 1"; do set -- $v
  DFSFM_ATTN_FOLD=$1 DFSFM_KV_EPI=$2 DFSFM_ENC_FUSED=$3 timeout 300 python bench.py --steps 5 --warmup 3 --skip-hp2 --skip-post --skip-img --skip-cpu > gpurun_out/r2a_bench_fold$1_kvepi$2_fused$3.json 2> gpurun_out/r2a_bench_fold$1_kvepi$2_fused$3.err
done
(timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err)
tail -c 600 gpurun_out/r2a_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"patch_conv11|kv_partial|attn_apply|fine_match_kernel|bicubic_merge|maxpool|SimEpi|stem_conv|KvEpi" -c 24 -o gpurun_out/r2a_hbm python tools/profile_step.py 1 256 > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_ncu.log
tail -12 gpurun_out/r2a_tests.log
