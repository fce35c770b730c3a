#!/bin/bash
# round-2 job F (2 GPUs): sharded configs over NCCL -- C4 pair graph with the sharded keypoint merge, C5 refinement chunks, the default line at N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
(timeout 500 $TR bench.py --gpus 2 --config c4 --c4-images 41 --steps 1 --warmup 1 > gpurun_out/r2f_c4_n2.json 2> gpurun_out/r2f_c4_n2.err); tail -c 400 gpurun_out/r2f_c4_n2.err
(timeout 300 python bench.py --gpus 1 --config c4 --c4-images 41 --steps 1 --warmup 1 > gpurun_out/r2f_c4_n1.json 2> gpurun_out/r2f_c4_n1.err); tail -c 300 gpurun_out/r2f_c4_n1.err
(timeout 300 $TR bench.py --gpus 2 --config c5 --c5-chunks 32 > gpurun_out/r2f_c5_n2.json 2> gpurun_out/r2f_c5_n2.err); tail -c 300 gpurun_out/r2f_c5_n2.err
(timeout 300 python bench.py --gpus 1 --config c5 --c5-chunks 32 > gpurun_out/r2f_c5_n1.json 2> gpurun_out/r2f_c5_n1.err); tail -c 300 gpurun_out/r2f_c5_n1.err
(timeout 500 $TR bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err); tail -c 300 gpurun_out/r2f_bench_n2.err
cat gpurun_out/r2f_c4_n2.json gpurun_out/r2f_c4_n1.json gpurun_out/r2f_c5_n2.json gpurun_out/r2f_c5_n1.json | cut -c 1-400
