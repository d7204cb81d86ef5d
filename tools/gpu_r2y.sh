#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/exp_r2q.py f32_rev70,f32_rev54,f32_rev100,f32_c1_3000,f32_c2_4000,f32_c2_4096,f32_c2_1000,f32_rev48,f32_c1_2000,f32_rev40 > gpurun_out/r2y_exp_f32_narrow.txt 2>&1; cat gpurun_out/r2y_exp_f32_narrow.txt
