#!/bin/bash
# round-2 job C: timeline of the encoder path, pair-worker A/B, ncu of the tensor-core kernels of the encoder / similarity path
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/r2c_tests.log
tail -5 gpurun_out/r2c_tests.log
DFSFM_ENC_FUSED=1 timeout 200 python tools/timeline.py 0 400 raw > gpurun_out/r2c_timeline_fused.log 2>&1
tail -3 gpurun_out/r2c_timeline_fused.log
for v in "1 1" "1 2" "0 2" "1 3"; do set -- $v
  DFSFM_ENC_FUSED=$1 DFSFM_BENCH_WORKERS=$2 timeout 300 python bench.py --steps 5 --warmup 3 --skip-hp2 --skip-post --skip-img --skip-cpu > gpurun_out/r2c_bench_fused$1_workers$2.json 2> gpurun_out/r2c_bench_fused$1_workers$2.err
  tail -c 300 gpurun_out/r2c_bench_fused$1_workers$2.err
done
DFSFM_ENC_FUSED=1 timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"enc256_fused|KvEpi|SimEpi|kvp_fold" -s 30 -c 14 -o gpurun_out/r2c_enc python tools/profile_step.py 1 0 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
