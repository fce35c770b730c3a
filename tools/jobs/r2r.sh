#!/bin/bash
# round-2 job R (1 GPU): KvEpi state reduction on warp-level tf32 MMAs: parity, timeline, bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_coarse_gpu.py tests/test_baseline_shapes_gpu.py tests/test_fullsize_gpu.py tests/test_engine_gpu.py -m gpu -q -x 2>&1 | tail -8) > gpurun_out/r2r_tests.log
tail -3 gpurun_out/r2r_tests.log
timeout 200 python tools/timeline.py 0 400 raw > gpurun_out/r2r_timeline.log 2>&1
B="timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --skip-cpu --skip-post --skip-img --skip-hp2"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d["roofline"]["kernel_ms_per_step"]; print(sys.argv[1], round(d["value"],1), round(d["e2e"]["value"],1), round(d.get("value_one_pair_in_flight",0),1), round(d.get("value_cached",0),1), k["conv"], k["enc_fused"], k["kvproj"], k["sim"])'
: > gpurun_out/r2r_spread.log
for i in 1 2 3; do $B 2>/dev/null | python -c "$P" kvmma_w3 >> gpurun_out/r2r_spread.log; done
cat gpurun_out/r2r_spread.log
