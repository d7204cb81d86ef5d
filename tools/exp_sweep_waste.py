"""Odd-size permutedims sweep under different padding-waste thresholds of the tile planner (tuning aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import strided_jl_b200 as sb
from strided_jl_b200.engine import Engine
from tools.sweep import col


def graph_time(eng, fn, reps=10):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        eng.set_stream(st.cuda_stream)
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(3):
            g.replay()
        e1.record(st)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * reps)


def main():
    for s in (25, 41, 54, 70, 91):
        shape = (s,) * 4
        n = s ** 4
        a = torch.randn(n, dtype=torch.float64, device="cuda")
        b = torch.empty_like(a)
        A, B = sb.StridedView(a, shape, col(shape)), sb.StridedView(b, shape, col(shape))
        for p in ((3, 2, 1, 0), (1, 2, 3, 0)):
            Ap = A.permutedims(p)
            row = []
            for env in ({}, {"SB_WASTE": "1.4"}, {"SB_WASTE": "1.6"}, {"SB_WASTE": "2.0"}, {"SB_NO_TMA": "1"}, {"SB_NO_TMA": "1", "SB_WASTE": "1.6"}):
                for k in list(os.environ):
                    if k.startswith("SB_"):
                        del os.environ[k]
                os.environ.update(env)
                eng = Engine(0)
                eng.set_sync(False)
                us = graph_time(eng, lambda: sb.run_mapreduce([], 0, 0, 0.0, shape, [B, Ap], engine=eng))
                pl = sb.plan_describe(sb.make_desc([], 0, 0, 0.0, shape, [B, Ap]))
                eng.close()
                row.append(f"{env or 'default'}: {us:7.2f}us {2 * 8 * n / us / 1e3:6.0f}GB/s tile={pl.get('tile')} tma={pl.get('tma')}")
            print(f"s={s} p={p}\n   " + "\n   ".join(row), flush=True)


if __name__ == "__main__":
    main()
