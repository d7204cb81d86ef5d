#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
SB_JIT_VERBOSE=1 python tools/exp_interp.py 2>&1 | tee gpurun_out/exp_interp_jit.txt | tail -16
ls ~/.cache/strided_b200 2>/dev/null | wc -l
