#!/bin/bash
mkdir -p gpurun_out
nvcc -std=c++17 -O3 --fmad=false -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_reduce tools/ubench_reduce.cu 2> gpurun_out/r2e_nvcc.log
timeout 120 /tmp/ubench_reduce 16777216 quick > gpurun_out/r2e_calib.txt 2>&1; cat gpurun_out/r2e_calib.txt
timeout 600 python tools/exp_r2e.py > gpurun_out/r2e_exp.txt 2>&1; cat gpurun_out/r2e_exp.txt
