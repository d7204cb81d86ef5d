#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
  for v in 0 1 2; do timeout 200 python tools/ab_lib.py tools/ab/lib_pf$v.so 100 >> gpurun_out/r2g_ab.txt 2>&1; done
  timeout 200 python tools/ab_lib.py tools/ab/libstrided_b200_r01.so 100 >> gpurun_out/r2g_ab.txt 2>&1
done
cat gpurun_out/r2g_ab.txt
timeout 300 python tools/exp_r2e.py > gpurun_out/r2g_exp.txt 2>&1; grep c5shard gpurun_out/r2g_exp.txt
