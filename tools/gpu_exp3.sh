#!/bin/bash
mkdir -p gpurun_out
( for s in 2 3; do SB_TMA_STAGES=$s timeout 120 python tools/time_case.py c2 300; done
  SB_NO_TMA=1 timeout 120 python tools/time_case.py c2 300
) 2>&1 | grep -E "us=|Error|error|Traceback" | tee gpurun_out/exp3.txt
