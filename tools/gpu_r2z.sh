#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/exp_reduce_dims.py > gpurun_out/r2z_reduce_dims_final.txt 2>&1; cat gpurun_out/r2z_reduce_dims_final.txt
