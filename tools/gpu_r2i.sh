#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
  timeout 300 python tools/ab_lib.py tools/ab/libstrided_b200_r01.so 50 >> gpurun_out/r2i_ab.txt 2>&1
  timeout 300 python tools/ab_lib.py strided.jl_b200/libstrided_b200.so 50 >> gpurun_out/r2i_ab.txt 2>&1
done
cat gpurun_out/r2i_ab.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2i_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -3 gpurun_out/r2i_bench.err
python - <<'PY'
import json
try:
    d = [json.loads(l) for l in open('gpurun_out/r2i_bench.json') if l.startswith('{')][-1]
    print({k: d[k] for k in ('value', 'ms_per_step')}, d['roofline']['frac'])
    for e in d.get('configs', []):
        print(f"{e['config']:78s} {e['ms']*1e3:8.2f} us {e['GBps']:8.1f} GB/s {e['frac_of_peak']:.3f}  {e.get('kernel','')}")
    print(json.dumps(d.get('sharded'), indent=1)[:900])
    print(d['e2e']['value'], d['e2e']['one_call_at_a_time'], d['e2e']['copy_only_ceiling'])
except Exception as e:
    print("bench parse failed", e)
PY
