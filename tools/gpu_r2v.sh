#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/exp_r2q.py c2_4000,c2_4096,c2_2048,c1_1024,c1_4096,c1_3000,rev48,rev96 > gpurun_out/r2v_exp_tma_vs_lsu2.txt 2>&1; cat gpurun_out/r2v_exp_tma_vs_lsu2.txt
