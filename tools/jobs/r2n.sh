#!/bin/bash
# round-2 job N (1 GPU): CLC probe (fixed), all GPU tests, the default line (3 pair workers, tile stealing, shuffle-based similarity pass)
mkdir -p gpurun_out
(nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/exp_clc tools/exp_clc.cu 2>/dev/null && timeout 60 /tmp/exp_clc) > gpurun_out/r2n_clc_probe.log 2>&1
cat gpurun_out/r2n_clc_probe.log
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/r2n_tests.log
tail -3 gpurun_out/r2n_tests.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
B="timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --skip-cpu --skip-post --skip-img --skip-hp2"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"],1), round(d["e2e"]["value"],1), round(d.get("value_one_pair_in_flight",0),1), round(d.get("value_cached",0),1), d["roofline"]["kernel_ms_per_step"]["conv"], d["roofline"]["kernel_ms_per_step"]["sim"])'
: > gpurun_out/r2n_spread.log
for i in 1 2; do $B 2>/dev/null | python -c "$P" steal_w3 >> gpurun_out/r2n_spread.log; done
for i in 1 2; do DFSFM_TILE_STEAL=0 $B 2>/dev/null | python -c "$P" static_w3 >> gpurun_out/r2n_spread.log; done
cat gpurun_out/r2n_spread.log
cut -c1-400 gpurun_out/r2n_bench.json
