#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/exp_r2q.py rev41,rev70,rev91,rev100,rot70,c2_3001,swap91 > gpurun_out/r2t_exp_ept.txt 2>&1; cat gpurun_out/r2t_exp_ept.txt
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/profile_batch.py 32 3 1 > gpurun_out/r02s_sanitizer_${tool}_group.log 2>&1; tail -2 gpurun_out/r02s_sanitizer_${tool}_group.log
done
