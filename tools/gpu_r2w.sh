#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/exp_r2q.py rev41,rev70,rev91,rev100,rot70,c2_3001,swap91,c1_1001,c3_32,c1_1000,rev64,rev128 > gpurun_out/r2w_exp_lsu_prefetch_l1.txt 2>&1; cat gpurun_out/r2w_exp_lsu_prefetch_l1.txt
timeout 900 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2w_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2w_bench.json') if l.startswith('{')][-1])
for e in d.get('configs', []):
    print(f"{e['config']:78s} {e['ms']*1e3:8.2f} us {e['GBps']:8.1f} GB/s {e['frac_of_peak']:.3f}  {e.get('kernel','')[:90]}")
PY
