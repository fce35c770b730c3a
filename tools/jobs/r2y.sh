#!/bin/bash
# round-2 job Y (1 GPU): compute-sanitizer memcheck over the coarse transformer (fused kernel with chunked hand-offs, KvEpi) and one small pair
mkdir -p gpurun_out
(timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python tests/check_transformer.py 2>&1 | tail -15) > gpurun_out/r2y_memcheck_transformer.log
tail -5 gpurun_out/r2y_memcheck_transformer.log
(timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15) > gpurun_out/r2y_memcheck_smoke.log
tail -5 gpurun_out/r2y_memcheck_smoke.log
