#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 200 --warmup 10 --extra > gpurun_out/bench.json 2> gpurun_out/bench_extra.log
cut -c1-330 gpurun_out/bench.json
python - <<'PY'
import json
for e in json.load(open('gpurun_out/bench_configs.json')):
    print(f"{e['config']:70s} {e['ms']*1e3:8.2f} us {e['GBps']:8.1f} GB/s {e['frac_of_peak']:.3f}")
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference_arm.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:map_tma -s 4 -c 1 -f -o gpurun_out/prof_c2_tma python tools/profile_case.py c2 6 > gpurun_out/ncu_c2.log 2>&1; tail -1 gpurun_out/ncu_c2.log
timeout 400 python tools/sweep.py > gpurun_out/sweep.txt 2>&1; tail -5 gpurun_out/sweep.txt
