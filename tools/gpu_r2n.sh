#!/bin/bash
mkdir -p gpurun_out
F="-std=c++17 -O3 --fmad=false --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -I strided.jl_b200/csrc"
nvcc $F -o /tmp/usp0 tools/ubench_stream_product.cu 2> gpurun_out/r2n_nvcc.log
nvcc $F -DSB_STREAM_CONSUMER_NOWAIT -o /tmp/usp1 tools/ubench_stream_product.cu 2>> gpurun_out/r2n_nvcc.log
for i in 1 2; do
  echo "## default" >> gpurun_out/r2n_stream_product.txt; timeout 120 /tmp/usp0 >> gpurun_out/r2n_stream_product.txt 2>&1
  echo "## consumers do not execute griddepcontrol.wait" >> gpurun_out/r2n_stream_product.txt; timeout 120 /tmp/usp1 >> gpurun_out/r2n_stream_product.txt 2>&1
done
cat gpurun_out/r2n_stream_product.txt
timeout 300 python tools/exp_r2e.py 2>&1 | grep c5shard | head -3
