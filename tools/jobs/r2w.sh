#!/bin/bash
# round-2 job W (2 GPUs): the default line at N=2 (torchrun) with the final kernels, and the C4 / C5 strong-scaling configs at N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29619"
(timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2w_bench_n2.json 2> gpurun_out/r2w_bench_n2.err); grep -v "UserWarning\|return func\|OMP_NUM\|\*\*\*\*" gpurun_out/r2w_bench_n2.err | tail -3
(timeout 600 $TR bench.py --gpus 2 --config c4 --c4-images 41 --steps 1 --warmup 1 > gpurun_out/r2w_c4_n2.json 2> gpurun_out/r2w_c4_n2.err); tail -c 300 gpurun_out/r2w_c4_n2.err
(timeout 600 $TR bench.py --gpus 2 --config c5 --c5-chunks 40 --steps 1 --warmup 1 > gpurun_out/r2w_c5_n2.json 2> gpurun_out/r2w_c5_n2.err); tail -c 300 gpurun_out/r2w_c5_n2.err
cut -c 1-260 gpurun_out/r2w_bench_n2.json; cut -c 1-300 gpurun_out/r2w_c4_n2.json; cut -c 1-300 gpurun_out/r2w_c5_n2.json
