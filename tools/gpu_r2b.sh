#!/bin/bash
# round 2, batch B: GPU suite, smoke, default bench line (configs + sharded records)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2b_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; tail -2 gpurun_out/r2b_smoke.log
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -3 gpurun_out/r2b_bench.err; cut -c1-400 gpurun_out/r2b_bench.json
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2b_bench.json'))
    for e in d.get('configs', []):
        print(f"{e['config']:70s} {e['ms']*1e3:8.2f} us {e['GBps']:8.1f} GB/s {e['frac_of_peak']:.3f}  {e.get('kernel','')}")
    print(json.dumps(d.get('sharded'), indent=1))
    print(json.dumps(d.get('e2e'), indent=1))
except Exception as e:
    print("bench parse failed", e)
PY
