#!/bin/bash
# round-2 job K (1 GPU): what makes the two-worker resident `value` bimodal?  PDL off, 3 and 4 workers
mkdir -p gpurun_out
B="timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --skip-cpu --skip-post --skip-img --skip-hp2"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"],1), round(d["e2e"]["value"],1), round(d.get("value_one_pair_in_flight",0),1))'
: > gpurun_out/r2k_spread.log
for i in 1 2 3 4; do DFSFM_PDL=0 $B 2>/dev/null | python -c "$P" pdl0_w2 >> gpurun_out/r2k_spread.log; done
for i in 1 2 3 4; do $B --workers-per-gpu 3 2>/dev/null | python -c "$P" w3 >> gpurun_out/r2k_spread.log; done
for i in 1 2; do $B --workers-per-gpu 4 2>/dev/null | python -c "$P" w4 >> gpurun_out/r2k_spread.log; done
for i in 1 2; do DFSFM_PDL=0 $B --workers-per-gpu 3 2>/dev/null | python -c "$P" pdl0_w3 >> gpurun_out/r2k_spread.log; done
cat gpurun_out/r2k_spread.log
