#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "rowsum or colsum or sum_dim or rowmax or sum_135 or initop" > gpurun_out/r02_san_memcheck_reductions.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02_san_memcheck_reductions.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "rowsum_768_float64 or colsum_768_float64 or sum_dim2" > gpurun_out/r02_san_racecheck_reductions.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02_san_racecheck_reductions.log
