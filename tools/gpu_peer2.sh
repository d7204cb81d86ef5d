#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 200 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q --tb=short -p no:cacheprovider 2>&1 | tail -12
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/c5_sharded.py 2> gpurun_out/c5_n$N.err | tee gpurun_out/c5_peer_n$N.json
tail -2 gpurun_out/c5_n$N.err
