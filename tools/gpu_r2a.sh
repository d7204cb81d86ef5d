#!/bin/bash
# round 2, batch A: reduction microbenchmark, plan variants, zero-copy end to end
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
nvcc -std=c++17 -O3 --fmad=false -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_reduce tools/ubench_reduce.cu 2> gpurun_out/r2a_nvcc.log
timeout 300 /tmp/ubench_reduce 16777216 > gpurun_out/r2a_ubench_reduce_2p24.txt 2>&1
timeout 120 /tmp/ubench_reduce 1048576 > gpurun_out/r2a_ubench_reduce_2p20.txt 2>&1
timeout 600 python tools/exp_r2a.py > gpurun_out/r2a_exp.txt 2>&1
tail -40 gpurun_out/r2a_ubench_reduce_2p24.txt
tail -30 gpurun_out/r2a_exp.txt
