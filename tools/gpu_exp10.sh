#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
SB_FORCE_EPT=8 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -p no:cacheprovider -k "device_pointers" 2>&1 | tail -2
( timeout 120 python tools/time_case.py c2 300
  SB_NO_TILE_DESC=1 timeout 120 python tools/time_case.py c2 300
  SB_DEBUG=nostore,noload timeout 120 python tools/time_case.py c2 300
  SB_TMA_STAGES=2 timeout 120 python tools/time_case.py c2 300
) 2>&1 | grep -E "us=|rror" | tee gpurun_out/exp10.txt
