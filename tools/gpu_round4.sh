#!/bin/bash
# full round: parity, sanitizer, bench (both arms), full ncu capture of the headline kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/profile_case.py all 1 > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/profile_case.py c2 1 > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck.log
timeout 600 python bench.py --steps 100 --warmup 5 --extra > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -2 gpurun_out/bench.err | cut -c1-2500; cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:map_tma -s 4 -c 1 -o gpurun_out/prof_c2_tma python tools/profile_case.py c2 6 > gpurun_out/ncu_c2.log 2>&1
tail -2 gpurun_out/ncu_c2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 60 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
