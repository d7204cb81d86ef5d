#!/bin/bash
mkdir -p gpurun_out
SB_JIT_SYNC=1 timeout 900 python tools/exp_probe.py 2>&1 | grep "v'\|README\|axpy" > gpurun_out/r2z_probe_jit_sync.txt; cat gpurun_out/r2z_probe_jit_sync.txt
