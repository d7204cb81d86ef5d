#!/bin/bash
# round 2, batch X (2 GPUs): sharded GPU test (NCCL + fused peer exchange + graph replay), bench.py --gpus 2 as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2x_gpus.txt
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r2x_pytest_sharded.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2x_pytest_sharded.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2x_bench_n2.json 2> gpurun_out/r2x_bench_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r2x_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2x_bench_n2.json') if l.startswith('{')][-1])
    print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step')})
    print(json.dumps(d.get('sharded'), indent=1)[:1800])
    print(json.dumps(d.get('e2e'), indent=1)[:600])
except Exception as e:
    print("bench parse failed", e)
PY
