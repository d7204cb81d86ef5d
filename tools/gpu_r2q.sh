#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/exp_r2q.py rev41,rev70,rev91,rev100,rot70,c2_3001 > gpurun_out/r2q_exp3.txt 2>&1; cat gpurun_out/r2q_exp3.txt
