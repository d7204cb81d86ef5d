#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_orbit.py -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_orbit.log 2>&1
tail -5 gpurun_out/pytest_orbit.log
timeout 600 python tools/exp_orbit.py 200 2>&1 | grep -E "us=|ERROR" | tee gpurun_out/exp_orbit.txt
timeout 900 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
