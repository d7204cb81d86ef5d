#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
  for l in libstrided_b200_r01 lib_10f0dc2 lib_c43e1d6 lib_pf0 lib_x_newtree_oldtma lib_y_oldtree_newtma; do timeout 200 python tools/ab_lib.py tools/ab/$l.so 100 2>&1 | grep c2 >> gpurun_out/r2h_ab.txt; done
done
cat gpurun_out/r2h_ab.txt
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_jit.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; tail -5 gpurun_out/r2h_pytest.log
