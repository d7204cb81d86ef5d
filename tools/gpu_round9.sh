#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_orbit.py -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_orbit.log 2>&1
tail -3 gpurun_out/pytest_orbit.log
timeout 600 python tools/exp_orbit.py 20 c2,c4,c4p 2>&1 | grep -E "us=|ERROR" | tee gpurun_out/exp_orbit.txt
timeout 900 python bench.py --steps 100 --warmup 5 --extra > gpurun_out/bench.json 2> gpurun_out/bench_extra.log
tail -c 3000 gpurun_out/bench.json
python - <<'PY'
import json
for e in json.load(open('gpurun_out/bench_configs.json')):
    print(f"{e['config']:70s} {e['ms']*1e3:8.2f} us {e['GBps']:8.1f} GB/s {e['frac_of_peak']:.3f}")
PY
