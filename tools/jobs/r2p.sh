#!/bin/bash
# round-2 job P (2 GPUs): full GPU suite on GPU 0, then the default line at N=2 (torchrun) and N=1
mkdir -p gpurun_out
(CUDA_VISIBLE_DEVICES=0 timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r2p_tests.log
tail -3 gpurun_out/r2p_tests.log
(CUDA_VISIBLE_DEVICES=0 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8) > gpurun_out/r2p_smoke.log
tail -6 gpurun_out/r2p_smoke.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617"
(timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2p_bench_n2.json 2> gpurun_out/r2p_bench_n2.err); grep -v "UserWarning\|return func\|OMP_NUM\|\*\*\*\*" gpurun_out/r2p_bench_n2.err | tail -3
(CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err); tail -c 300 gpurun_out/r2p_bench_n1.err
cut -c 1-260 gpurun_out/r2p_bench_n2.json; cut -c 1-260 gpurun_out/r2p_bench_n1.json
