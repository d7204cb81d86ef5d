#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/last_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/last_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/last_smoke.log
timeout 600 python tools/exp_probe_reduce.py 2>&1 | grep "maximum\|dot\|sum(A) dense" > gpurun_out/r2z_probe_reduce_functors.txt; cat gpurun_out/r2z_probe_reduce_functors.txt
