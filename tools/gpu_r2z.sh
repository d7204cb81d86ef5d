#!/bin/bash
mkdir -p gpurun_out
SB_JIT_SYNC=1 timeout 600 python tools/exp_probe_reduce.py 2>&1 | grep "maximum\|dot" > gpurun_out/r2z_probe_reduce_jit.txt; cat gpurun_out/r2z_probe_reduce_jit.txt
