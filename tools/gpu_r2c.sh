#!/bin/bash
# round 2, batch C (2 GPUs): sharded GPU test (NCCL + fused peer exchange + graph replay), bench.py --gpus 2 as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c_gpus.txt
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r2c_pytest_sharded.log 2>&1; tail -5 gpurun_out/r2c_pytest_sharded.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err; tail -5 gpurun_out/r2c_bench_n2.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2c_bench_n2.json'))
    print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step')})
    print(json.dumps(d.get('sharded'), indent=1))
    print(json.dumps(d.get('e2e'), indent=1))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python bench.py --steps 20 --warmup 3 --no-configs --no-sharded > gpurun_out/r2c_bench_n1_short.json 2> gpurun_out/r2c_bench_n1_short.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2c_bench_n1_short.json'))
    print(json.dumps(d.get('e2e'), indent=1))
except Exception as e:
    print("bench parse failed", e)
PY
