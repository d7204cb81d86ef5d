#!/bin/bash
mkdir -p gpurun_out
( for tb in "6,5" "5,6" "7,4" "4,7" "8,3"; do SB_TILE_BITS=$tb timeout 120 python tools/time_case.py c2 300; done
  for l2 in 0 64 256; do SB_TMA_L2PROMO=$l2 timeout 120 python tools/time_case.py c2 300; done
  SB_NO_TMA=1 SB_TILE_BITS="5,6" timeout 120 python tools/time_case.py c2 300
  SB_NO_TMA=1 SB_TILE_BITS="7,4" timeout 120 python tools/time_case.py c2 300
) 2>&1 | grep -E "us=|rror" | tee gpurun_out/exp6.txt
