#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/exp_r2q.py f32_rev54,rev55,rev59,rev27,c1_119,f32_c1_955,rev41,rev70,c3_32,c1_1000 > gpurun_out/r2z_exp_auto_balanced.txt 2>&1; cat gpurun_out/r2z_exp_auto_balanced.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2z_pytest_gpu.log
timeout 600 python tools/sweep.py > gpurun_out/r2z_sweep2.txt 2>&1; head -16 gpurun_out/r2z_sweep2.txt
