#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
SB_FORCE_EPT=8 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -p no:cacheprovider -k "device_pointers" > gpurun_out/pytest_gpu_ept8.log 2>&1
tail -2 gpurun_out/pytest_gpu_ept8.log
( for cfg in c2 c1 c3 c4 c4p c5 c5shard; do timeout 120 python tools/time_case.py $cfg 200; done
  SB_NO_VEC=1 timeout 120 python tools/time_case.py c2 200
  SB_NO_VEC=1 timeout 120 python tools/time_case.py c4 200
  SB_NO_TMA=1 timeout 120 python tools/time_case.py c2 200
  SB_FORCE_EPT=8 timeout 120 python tools/time_case.py c4 200
) 2>&1 | grep -E "us=|Error|error|Traceback" | tee gpurun_out/exp4.txt
timeout 120 python tools/exp_copy.py 2>&1 | tee gpurun_out/exp_copy.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__grid_size
timeout 600 ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/metrics.csv python tools/profile_case.py all 3 > gpurun_out/ncu_metrics.log 2>&1
tail -1 gpurun_out/ncu_metrics.log
