#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python tools/time_case.py c2 300
  SB_DEBUG=nostore timeout 120 python tools/time_case.py c2 300
  SB_DEBUG=noload timeout 120 python tools/time_case.py c2 300
  SB_DEBUG=nostore,noload timeout 120 python tools/time_case.py c2 300
) 2>&1 | grep -E "us=|rror" | tee gpurun_out/exp9.txt
