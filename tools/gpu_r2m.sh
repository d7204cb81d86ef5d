#!/bin/bash
# round 2, batch M (8 GPUs): bench.py --gpus 8 exactly as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2m_ngpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2m_bench_n8.json 2> gpurun_out/r2m_bench_n8.err; tail -3 gpurun_out/r2m_bench_n8.err
grep "^{" gpurun_out/r2m_bench_n8.json | tail -1 | cut -c1-300
