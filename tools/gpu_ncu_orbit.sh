#!/bin/bash
mkdir -p gpurun_out
for cfg in c4 c2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:map_orbit -s 4 -c 1 -f -o gpurun_out/prof_orbit_$cfg python tools/profile_case.py $cfg 6 > gpurun_out/ncu_$cfg.log 2>&1
  tail -2 gpurun_out/ncu_$cfg.log
done
ls -la gpurun_out/*.ncu-rep
