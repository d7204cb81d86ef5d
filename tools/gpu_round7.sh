#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( for cfg in c5 c5shard c2; do timeout 120 python tools/time_case.py $cfg 200; done ) 2>&1 | grep -E "us=|rror" | tee gpurun_out/exp8.txt
python tools/c5_sharded.py | tee gpurun_out/c5_n1.json
timeout 300 python tools/sweep.py 2>&1 | tail -16 | tee gpurun_out/sweep_sum.txt
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/profile_case.py c5shard 1 > gpurun_out/sanitizer_racecheck_reduce.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/sanitizer_racecheck_reduce.log
