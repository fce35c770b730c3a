#!/bin/bash
# round-2 job D: KvEpi 4x4 tiles, TMA-store epilogue of the fused kernel, pair workers
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/r2d_tests.log
tail -5 gpurun_out/r2d_tests.log
timeout 200 python tools/timeline.py 0 400 raw > gpurun_out/r2d_timeline.log 2>&1
tail -2 gpurun_out/r2d_timeline.log
for w in 1 2 3; do
  DFSFM_BENCH_WORKERS=$w timeout 300 python bench.py --steps 5 --warmup 3 --skip-hp2 --skip-post --skip-img --skip-cpu > gpurun_out/r2d_bench_workers$w.json 2> gpurun_out/r2d_bench_workers$w.err
  tail -c 300 gpurun_out/r2d_bench_workers$w.err
done
