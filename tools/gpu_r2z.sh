#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2z_smoke.log
