#!/bin/bash
# round-2 job U (1 GPU): elect.sync for the single-thread TMA / MMA roles (UTCHMMA / UTMALDG without the per-instruction warp loop): tests, bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r2u_tests.log
tail -3 gpurun_out/r2u_tests.log
B="timeout 200 python bench.py --gpus 1 --steps 5 --warmup 3 --skip-cpu --skip-post --skip-img"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d["roofline"]["kernel_ms_per_step"]; h=d.get("hp2",{}); hk=h.get("roofline",{}).get("kernel_ms_per_chunk",{}); print(sys.argv[1], round(d["value"],1), round(d["e2e"]["value"],1), round(d.get("value_one_pair_in_flight",0),1), round(d.get("value_cached",0),1), k["conv"], k["enc_fused"], k["kvproj"], k["sim"], round(h.get("value",0)), hk.get("pconv"), hk.get("mlp_fused"), hk.get("lin"), d["clocks"]["sm_mhz"])'
: > gpurun_out/r2u_spread.log
for i in 1 2 3; do $B 2>/dev/null | python -c "$P" elect_w3 >> gpurun_out/r2u_spread.log; done
cat gpurun_out/r2u_spread.log
timeout 200 python tools/timeline.py 0 60 raw > gpurun_out/r2u_timeline.log 2>&1
