#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/last_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/last_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/last_smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-configs --no-sharded 2>/dev/null | tail -1 | cut -c1-260
