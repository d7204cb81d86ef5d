#!/bin/bash
# round 2, batch S: ncu evidence for the kernels added in this session (exported as CSV on the box), sanitizers on them
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02s_launches_bench.csv python bench.py --steps 20 --warmup 3 > gpurun_out/r02s_bench_under_ncu.log 2>&1
cap() { # name regex command...
  local name=$1 rx=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o /tmp/prof_$name "$@" > gpurun_out/r02s_ncu_$name.log 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/r02s_ncu_${name}_raw.csv 2>/dev/null
  ls -la /tmp/prof_$name.ncu-rep gpurun_out/r02s_ncu_${name}_raw.csv
}
cap stream_inter reduce_stream python tools/profile_case.py c5 6
cap tma_group map_tma_group python tools/profile_batch.py 32 8 6
cap rev91_lsu map_tile python tools/profile_case.py rev91 6
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/profile_batch.py 16 5 2 > gpurun_out/r02s_sanitizer_${tool}_group.log 2>&1; tail -2 gpurun_out/r02s_sanitizer_${tool}_group.log
  timeout 600 compute-sanitizer --tool $tool python tools/profile_case.py c5small 2 > gpurun_out/r02s_sanitizer_${tool}_stream_inter.log 2>&1; tail -2 gpurun_out/r02s_sanitizer_${tool}_stream_inter.log
done
du -sh gpurun_out
