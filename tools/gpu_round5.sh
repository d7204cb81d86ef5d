#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/sweep.py 2>&1 | tee gpurun_out/sweep.txt | tail -34
( for cfg in c2 c1 c3 c4 c4p c5 c5shard; do timeout 120 python tools/time_case.py $cfg 200; done ) 2>&1 | grep -E "us=|rror" | tee gpurun_out/exp5.txt
