"""Config 5 as specified: Float64 8 x 4096 x 4096 `mapreduce(abs2, +, A; dims=(2,3))` sharded on dim 1 over the ranks.
Rank g holds the DENSE slab A[g,:,:] (4096x4096, 128 MiB) in its own HBM and produces out[g]: no collective.
Also: the complete reduction sum(abs2, A) with ONE all-reduce of one element.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/c5_sharded.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import strided_jl_b200 as sb
from strided_jl_b200 import sharded


def _peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


PEAK = _peak()


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    G, K = 8, 4096
    per = G // world if world <= G else 1
    torch.manual_seed(1234 + rank)
    slab = torch.randn(per * K * K, dtype=torch.float64, device=dev)  # `per` slices, each stored densely
    A = sb.StridedView(slab, (per, K, K), (K * K, 1, K))              # A_local[g, j, k]
    out = torch.zeros(per, dtype=torch.float64, device=dev)
    O = sb.StridedView(out, (per, K, K), (1, 0, 0))
    eng = sb.get_engine(local)
    eng.set_sync(False)
    prog = [(0, 0, 0.0, 0.0), (2, sb.abi.FN["abs2"], 0.0, 0.0)]

    def step():
        sb.run_mapreduce(prog, 1, 1, 0.0, (per, K, K), [O, A])  # initop=zero: out = sum abs2

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    want = (slab.view(per, K * K) ** 2).sum(dim=1)
    assert torch.allclose(out, want, rtol=1e-12), (out, want)
    steps = 50
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # device time of `steps` back-to-back launches: one CUDA graph, replayed (no host launch cost in the number)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        step()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(steps):
                step()
        g.replay()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        g.replay()
        e1.record(side)
        torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    # complete reduction with the single all-reduce
    tot = sharded.sharded_mapreduce("abs2", "+", A, shard_dim=0)
    f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    f0.record()
    for _ in range(10):
        tot = sharded.sharded_mapreduce("abs2", "+", A, shard_dim=0)
    f1.record()
    torch.cuda.synchronize()
    ms_full = torch.tensor([f0.elapsed_time(f1) / 10], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_full, op=dist.ReduceOp.MAX)
    if rank == 0:
        bytes_total = G * K * K * 8 if world <= G else world * K * K * 8
        print(json.dumps({"config": "C5 f64 8x4096x4096 mapreduce(abs2,+;dims=(2,3)) sharded on dim 1", "n_gpus": world,
                          "slices_per_gpu": per, "us_per_step_max_over_ranks": ms.item() * 1e3,
                          "aggregate_GBps": bytes_total / (ms.item() * 1e-3) / 1e9,
                          "frac_of_N_x_peak": bytes_total / (ms.item() * 1e-3) / 1e9 / (PEAK * world),
                          "complete_reduction_with_one_allreduce_us (incl. host sync + .item())": ms_full.item() * 1e3,
                          "placement": "dense 4096x4096 slab per slice in each GPU's HBM; no data-path collective"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
