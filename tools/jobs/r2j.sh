#!/bin/bash
# round-2 job J (1 GPU): schedule A/B subprocess test, fine_match back to the old loop, and the run-to-run spread of the two-worker `value`
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_coarse_gpu.py tests/test_refine_gpu.py -m gpu -q 2>&1 | tail -8) > gpurun_out/r2j_tests.log
tail -3 gpurun_out/r2j_tests.log
B="timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --skip-cpu --skip-post --skip-img"
for i in 1 2 3; do $B --skip-hp2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default2', d['value'], d['e2e']['value'], d.get('value_one_pair_in_flight'))"; done > gpurun_out/r2j_spread.log 2>&1
for i in 1 2 3; do DFSFM_BENCH_STAGGER_MS=2.2 $B --skip-hp2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stagger2.2', d['value'], d['e2e']['value'])"; done >> gpurun_out/r2j_spread.log 2>&1
for i in 1 2; do $B --skip-hp2 --workers-per-gpu 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('workers3', d['value'], d['e2e']['value'])"; done >> gpurun_out/r2j_spread.log 2>&1
$B > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
cat gpurun_out/r2j_spread.log
