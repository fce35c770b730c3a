#!/bin/bash
# round-2 job G: validate the latest fused-kernel changes (early x part, weight prefetch ahead of PDL wait), patch_conv11 4-pixel; timeline; bench; post/image ncu
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/r2g_tests.log
tail -5 gpurun_out/r2g_tests.log
timeout 200 python tools/timeline.py 0 400 raw > gpurun_out/r2g_timeline.log 2>&1
(timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err); tail -c 300 gpurun_out/r2g_bench.err
timeout 200 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"rs_scatter|lanczos_h|lanczos_v|patch_conv11" -s 13 -c 16 -o gpurun_out/r2g_post python tools/profile_post.py > gpurun_out/r2g_ncu_post.log 2>&1
tail -2 gpurun_out/r2g_ncu_post.log
