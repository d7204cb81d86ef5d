#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_orbit.py -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_orbit.log 2>&1
tail -3 gpurun_out/pytest_orbit.log
SB_ORBIT_TMASTORE=1 timeout 600 python -m pytest tests/test_orbit.py -m gpu -x -q --tb=short -p no:cacheprovider 2>&1 | tail -1
timeout 600 python tools/exp_orbit.py 20 c2,c4,c4p,c1,c3 2>&1 | grep -E "us=|ERROR" | tee gpurun_out/exp_orbit.txt
