#!/bin/bash
# first GPU visit: parity, bench, launch list, one full ncu capture of the C2 kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -40 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 --extra > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python tools/profile_case.py all 6 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:map_tile -s 4 -c 2 -o gpurun_out/prof_c2 python tools/profile_case.py c2 6 > gpurun_out/ncu_c2.log 2>&1
tail -3 gpurun_out/ncu_c2.log
ls -la gpurun_out
