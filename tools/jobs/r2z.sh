#!/bin/bash
# round-2 job Z (1 GPU): 208- / 256-channel convolutions as two column tiles (double-buffered accumulators + slab): parity, bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/r2z_tests.log
tail -3 gpurun_out/r2z_tests.log
B="timeout 200 python bench.py --gpus 1 --steps 5 --warmup 3 --skip-cpu --skip-post --skip-img"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d["roofline"]["kernel_ms_per_step"]; h=d.get("hp2",{}); hk=h.get("roofline",{}).get("kernel_ms_per_chunk",{}); print(sys.argv[1], round(d["value"],1), round(d["e2e"]["value"],1), round(d.get("value_one_pair_in_flight",0),1), round(d.get("value_cached",0),1), k["conv"], k["enc_fused"], k["kvproj"], k["sim"], round(h.get("value",0)), hk.get("pconv"), d["roofline"]["frac"], d["clocks"]["sm_mhz"])'
: > gpurun_out/r2z_spread.log
for i in 1 2; do $B 2>/dev/null | python -c "$P" ntile_w3 >> gpurun_out/r2z_spread.log; done
cat gpurun_out/r2z_spread.log
