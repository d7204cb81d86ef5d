#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/exp_r2q.py rev41,rev70,rev91,rev100,rot70,c2_3001,swap91,c1_1001 > gpurun_out/r2u_exp_lsu_desc2.txt 2>&1; cat gpurun_out/r2u_exp_lsu_desc2.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2u_pytest_gpu.log
