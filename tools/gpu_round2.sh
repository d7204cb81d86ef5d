#!/bin/bash
# parity (fail fast), bench with all configs, ncu metric sweep over every BASELINE config
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 --extra > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err | cut -c1-3000; cat gpurun_out/bench.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__grid_size
timeout 600 ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/metrics.csv python tools/profile_case.py all 3 > gpurun_out/ncu_metrics.log 2>&1
tail -2 gpurun_out/ncu_metrics.log
