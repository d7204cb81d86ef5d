#!/bin/bash
mkdir -p gpurun_out
python tools/abi_latency.py 2>&1 | tee gpurun_out/abi_latency.txt
python tools/exp_interp.py 2>&1 | tee gpurun_out/exp_interp.txt
