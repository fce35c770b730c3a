#!/bin/bash
# round-2 job H2 (2 GPUs): the default line at N=2 (weak scaling, two pair workers per rank) after the thread-safety fix
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612"
(timeout 500 $TR bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err); grep -v "UserWarning\|return func" gpurun_out/r2h_bench_n2.err | tail -5
cut -c 1-600 gpurun_out/r2h_bench_n2.json
