#!/bin/bash
# compute-sanitizer memcheck over the tests that exercise the kernels changed in the second half of round 2
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py -x -q -k "batch or group or odd or stream_inter or plan_table" > gpurun_out/r02_san_memcheck_tests.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_san_memcheck_tests.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_batch.py -x -q -k "same_plan or grouped" > gpurun_out/r02_san_racecheck_group_tests.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_san_racecheck_group_tests.log
