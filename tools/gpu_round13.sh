#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/exp_orbit.py 20 c1,c3,c4p,c5,c5shard 2>&1 | grep -E "us=|ERROR" | tee gpurun_out/exp_pdl.txt
python tools/c5_sharded.py | tee gpurun_out/c5_n1.json
