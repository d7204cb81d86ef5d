#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/r2x_bench_n4.json 2> gpurun_out/r2x_bench_n4.err; echo "rc=$?"; tail -2 gpurun_out/r2x_bench_n4.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2x_bench_n4.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step')})
s = d['sharded']; print(s['us_per_step_max_over_ranks'], s['frac_of_N_x_peak'], s.get('strong_scaling_efficiency'), s['complete_reduction']['fused_peer_exchange_us'], s['complete_reduction']['local_kernel_plus_nccl_allreduce_eager_us'])
PY
