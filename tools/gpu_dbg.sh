#!/bin/bash
mkdir -p gpurun_out
{
python tools/dbg_stream.py sum_100_100_2_float64 stream_dims12_float64 stream_inter8_float64 stream_inter2_float64 c5_8_256
SB_NO_STREAM_INTER=1 python tools/dbg_stream.py sum_100_100_2_float64 c5_8_256
SB_NO_STREAM=1 python tools/dbg_stream.py sum_100_100_2_float64
timeout 300 compute-sanitizer --tool memcheck python tools/dbg_stream.py sum_100_100_2_float64 2>&1 | tail -30
} > gpurun_out/dbg_stream.txt 2>&1
cat gpurun_out/dbg_stream.txt
