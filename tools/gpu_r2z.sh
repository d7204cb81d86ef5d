#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2z_smoke.log
timeout 900 python tools/exp_r2q.py f32_rev64,f32_c1_4096,f32_c2_4000,f32_c1_1000,f32_rot64,c2_4000,rev64 > gpurun_out/r2z_exp_f32_default.txt 2>&1; cat gpurun_out/r2z_exp_f32_default.txt
timeout 600 python tools/profile_case.py c4 3 > gpurun_out/r2z_c4.log 2>&1; tail -1 gpurun_out/r2z_c4.log
