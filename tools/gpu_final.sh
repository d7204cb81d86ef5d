#!/bin/bash
# final batch of the round: what the driver runs (GPU suite, smoke, both bench arms), outputs kept under gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/final_bench_reference.json | tail -1
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/final_bench.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('one_call_at_a_time'), d['clocks'])
for e in d.get('configs', []):
    print(f"{e['config']:78s} {e['ms']*1e3:8.2f} us {e['GBps']:8.1f} GB/s {e['frac_of_peak']:.3f}")
PY
