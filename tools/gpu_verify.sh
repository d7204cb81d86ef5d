#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
# ncu launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_bench.csv')) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    try:
        agg[r[ki][:90]].append(float(r[vi].replace(',', '')))
    except ValueError:
        pass
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{len(v):4d} x {sum(v)/len(v)/1e3:9.2f} us  {100*sum(v)/tot:5.1f}%  {k}")
PY
# sanitizers on the orbit kernel (small shapes: TMA load + TMA store, edge tiles, two-view variant)
for cfg in c4s c4e; do
  timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/profile_case.py $cfg 1 > gpurun_out/sanitizer_memcheck_orbit_$cfg.log 2>&1; echo "memcheck $cfg rc=$?"; tail -1 gpurun_out/sanitizer_memcheck_orbit_$cfg.log
done
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/profile_case.py c4s 1 > gpurun_out/sanitizer_racecheck_orbit.log 2>&1; echo "racecheck rc=$?"; tail -1 gpurun_out/sanitizer_racecheck_orbit.log
SB_ORBIT_NIN2=1 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/profile_case.py c2s 1 > gpurun_out/sanitizer_memcheck_orbit_c2s.log 2>&1; echo "memcheck c2s rc=$?"; tail -1 gpurun_out/sanitizer_memcheck_orbit_c2s.log
