#!/bin/bash
# round-2 job V (1 GPU): the final default line (with the CPU baselines), the reference arm, the ncu launch list of the same command and
# ncu --set full of the final conv / fused / KvEpi / similarity kernels
mkdir -p gpurun_out
(timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err); tail -c 300 gpurun_out/r2v_bench.err
(timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2v_bench_reference.json 2> gpurun_out/r2v_bench_reference.err); tail -c 300 gpurun_out/r2v_bench_reference.err
DFSFM_BENCH_WORKERS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 1 --warmup 3 --skip-hp2 --skip-post --skip-img --skip-cpu > gpurun_out/r2v_launches_bench.log 2>&1
tail -c 200 gpurun_out/r2v_launches_bench.log
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"enc256_fused|KvEpi|ConvEpi|SimEpi" -s 14 -c 16 -o gpurun_out/r2v_final python tools/profile_step.py 1 0 > gpurun_out/r2v_ncu.log 2>&1
tail -2 gpurun_out/r2v_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"enc256_fused|KvEpi|SimEpi" -s 30 -c 12 -o gpurun_out/r2v_enc python tools/profile_step.py 1 0 > gpurun_out/r2v_ncu2.log 2>&1
tail -2 gpurun_out/r2v_ncu2.log
ls -la gpurun_out/r2v_*.ncu-rep
cut -c1-300 gpurun_out/r2v_bench.json
