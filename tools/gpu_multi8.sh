#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/c5_sharded.py 2> gpurun_out/c5_n$N.err | tee gpurun_out/c5_n$N.json
tail -2 gpurun_out/c5_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err; cut -c1-300 gpurun_out/bench_n$N.json
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q --tb=short -p no:cacheprovider 2>&1 | tail -2
