#!/bin/bash
# round-2 job E: the headline bench line (2 pair workers), launch list of the same command under ncu, ncu --set full of the final encoder kernels
mkdir -p gpurun_out
(timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err); tail -c 400 gpurun_out/r2e_bench.err
(timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2e_bench_reference.json 2> gpurun_out/r2e_bench_reference.err); tail -c 300 gpurun_out/r2e_bench_reference.err
DFSFM_BENCH_WORKERS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 1 --warmup 3 --skip-hp2 --skip-post --skip-img --skip-cpu > gpurun_out/r2e_launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"enc256_fused|KvEpi|kvp_fold|ConvEpi" -s 40 -c 10 -o gpurun_out/r2e_enc python tools/profile_step.py 1 0 > gpurun_out/r2e_ncu.log 2>&1
tail -2 gpurun_out/r2e_ncu.log
timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"rs_scatter|lanczos_h|lanczos_v" -c 6 -o gpurun_out/r2e_post python bench.py --steps 1 --warmup 3 --skip-hp2 --skip-cpu > gpurun_out/r2e_ncu_post.log 2>&1
tail -2 gpurun_out/r2e_ncu_post.log
