#!/bin/bash
# round 2, batch K: ncu evidence, exported as CSV on the box (the .ncu-rep files exceed the 64 MiB that gpurun brings back)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1
cap() { # name regex case
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 4 -c 1 -f -o /tmp/prof_$1 python tools/profile_case.py $3 6 > gpurun_out/r02_ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details --csv > gpurun_out/r02_ncu_$1_details.csv 2>/dev/null
  ls -la /tmp/prof_$1.ncu-rep gpurun_out/r02_ncu_$1_raw.csv
}
cap stream reduce_stream c5shard
cap c2_tma map_tma c2
cap c4_orbit map_orbit c4
cap c3_tma map_tma c3
cap c4p_orbit map_orbit c4p
cap c5_reduce reduce_tile c5
cp /tmp/prof_c2_tma.ncu-rep gpurun_out/r02_prof_c2_tma.ncu-rep
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_k_smoke.log 2>&1; tail -1 gpurun_out/r02_k_smoke.log
timeout 300 python tools/exp_r2j.py > gpurun_out/r02_k_exp.txt 2>&1; grep -E "env=\{\} " gpurun_out/r02_k_exp.txt
du -sh gpurun_out
