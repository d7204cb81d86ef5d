#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/exp_reduce_dims.py > gpurun_out/r2z_reduce_dims_final.txt 2>&1; cat gpurun_out/r2z_reduce_dims_final.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2z_pytest_gpu.log
timeout 600 python tools/time_case.py c5 50; SB_NO_STREAM_INTER=1 timeout 600 python tools/time_case.py c5 50
