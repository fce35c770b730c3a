#!/bin/bash
# round-2 job L (1 GPU): pair-worker pool test + backbone baton A/B (run-to-run spread of the two- and three-worker `value`)
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_stage_gpu.py -m gpu -q 2>&1 | tail -8) > gpurun_out/r2l_tests.log
tail -3 gpurun_out/r2l_tests.log
B="timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --skip-cpu --skip-post --skip-img --skip-hp2"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"],1), round(d["e2e"]["value"],1), round(d.get("value_one_pair_in_flight",0),1), round(d.get("value_cached",0),1))'
: > gpurun_out/r2l_spread.log
for i in 1 2 3 4; do $B 2>/dev/null | python -c "$P" baton_w2 >> gpurun_out/r2l_spread.log; done
for i in 1 2; do $B --workers-per-gpu 3 2>/dev/null | python -c "$P" baton_w3 >> gpurun_out/r2l_spread.log; done
for i in 1 2; do DFSFM_BACKBONE_BATON=0 $B 2>/dev/null | python -c "$P" nobaton_w2 >> gpurun_out/r2l_spread.log; done
cat gpurun_out/r2l_spread.log
