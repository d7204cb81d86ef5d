#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/c5_sharded.py 2> gpurun_out/c5_n$N.err | tee gpurun_out/c5_n$N.json
tail -2 gpurun_out/c5_n$N.err
