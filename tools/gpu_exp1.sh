#!/bin/bash
mkdir -p gpurun_out
for cfg in c2 c1 c3 c4 c4p; do
  for e in 4 8 16; do
    SB_FORCE_EPT=$e timeout 120 python tools/time_case.py $cfg 200 2>&1 | tail -1
  done
done | tee gpurun_out/exp1.txt
timeout 120 python tools/time_case.py c5 50 | tail -1 | tee -a gpurun_out/exp1.txt
timeout 120 python tools/time_case.py c5shard 100 | tail -1 | tee -a gpurun_out/exp1.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:map_tile -s 4 -c 1 -o gpurun_out/prof_c2_v2 python tools/profile_case.py c2 6 > gpurun_out/ncu_c2.log 2>&1
tail -2 gpurun_out/ncu_c2.log
