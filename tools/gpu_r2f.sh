#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
  timeout 200 python tools/ab_lib.py tools/ab/libstrided_b200_r01.so 100 >> gpurun_out/r2f_ab.txt 2>&1
  timeout 200 python tools/ab_lib.py strided.jl_b200/libstrided_b200.so 100 >> gpurun_out/r2f_ab.txt 2>&1
done
cat gpurun_out/r2f_ab.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2f_pytest_gpu.log
