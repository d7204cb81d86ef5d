#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/exp_r2q.py rot70,rot100,t5001,t3001,rev91 > gpurun_out/r2z_exp_hot_order3.txt 2>&1; cat gpurun_out/r2z_exp_hot_order3.txt
timeout 600 python tools/sweep.py > gpurun_out/r2z_sweep.txt 2>&1; head -18 gpurun_out/r2z_sweep.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest_gpu.log
