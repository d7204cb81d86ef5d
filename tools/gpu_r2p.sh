#!/bin/bash
# round 2, batch P: GPU suite (exit code matters: no crash at interpreter exit), smoke, bench line, orbit split experiment
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2p_pytest_gpu.log; tail -4 gpurun_out/r2p_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2p_smoke.log
timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2p_bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2p_bench.json') if l.startswith('{')][-1])
    for e in d.get('configs', []):
        print(f"{e['config']:78s} {e['ms']*1e3:8.2f} us {e['GBps']:8.1f} GB/s {e['frac_of_peak']:.3f}  {e.get('kernel','')[:90]}")
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python tools/exp_r2p.py > gpurun_out/r2p_exp.txt 2>&1; cat gpurun_out/r2p_exp.txt
