#!/bin/bash
# round 2, batch J: ncu evidence (launch list of bench.py, full captures of the stream / TMA / orbit kernels), TMA-vs-LSU on 128-byte rows
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02_launches_bench.csv')) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    try:
        agg[r[ki][:100]].append(float(r[vi].replace(',', '')))
    except ValueError:
        pass
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{len(v):4d} x {sum(v)/len(v)/1e3:9.2f} us  {100*sum(v)/tot:5.1f}%  {k}")
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:reduce_stream -s 4 -c 1 -f -o gpurun_out/r02_prof_stream python tools/profile_case.py c5shard 6 > gpurun_out/r02_ncu_stream.log 2>&1; tail -1 gpurun_out/r02_ncu_stream.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:map_tma -s 4 -c 1 -f -o gpurun_out/r02_prof_c2_tma python tools/profile_case.py c2 6 > gpurun_out/r02_ncu_c2.log 2>&1; tail -1 gpurun_out/r02_ncu_c2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:map_orbit -s 4 -c 1 -f -o gpurun_out/r02_prof_c4_orbit python tools/profile_case.py c4 6 > gpurun_out/r02_ncu_c4.log 2>&1; tail -1 gpurun_out/r02_ncu_c4.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:map_tma -s 4 -c 1 -f -o gpurun_out/r02_prof_c3_tma python tools/profile_case.py c3 6 > gpurun_out/r02_ncu_c3.log 2>&1; tail -1 gpurun_out/r02_ncu_c3.log
timeout 600 python tools/exp_r2j.py > gpurun_out/r02_j_exp.txt 2>&1; cat gpurun_out/r02_j_exp.txt
ls -la gpurun_out/*.ncu-rep
