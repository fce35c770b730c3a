#!/bin/bash
# round-2 job I (2 GPUs): N=1 tests (fine_match + similarity changes) then the default line at N=2 and N=1 on the same box
mkdir -p gpurun_out
(CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -20) > gpurun_out/r2i_tests.log
tail -4 gpurun_out/r2i_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613"
(timeout 500 $TR bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err); grep -v "UserWarning\|return func" gpurun_out/r2i_bench_n2.err | tail -3
(CUDA_VISIBLE_DEVICES=0 timeout 500 python bench.py --gpus 1 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err); tail -c 200 gpurun_out/r2i_bench_n1.err
cut -c 1-300 gpurun_out/r2i_bench_n2.json; cut -c 1-300 gpurun_out/r2i_bench_n1.json
