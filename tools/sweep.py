"""Size sweep mirroring the reference's own benchmark harness (benchmarks/benchtests.jl:9-42):
`benchmark_permute` for p = (4,3,2,1), (2,3,4,1), (3,4,1,2) on Float64 s^4 arrays and `benchmark_sum` on 1-D arrays,
sizes ceil(2^(2:1.5:20)) elements in total (per-dim size = round(total^(1/4))).  Device time per call is measured by
replaying a CUDA graph of REPS launches (no host overhead), so the small-size end shows the launch/latency floor
-- the GPU analog of the reference's MINTHREADLENGTH crossover (src/mapreduce.jl:141).

    python tools/sweep.py > profiles/sweep.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import strided_jl_b200 as sb

PEAK = 6494.9


def graph_time(fn, reps=20, outer=5):
    """us per call, device time: capture `reps` launches into a CUDA graph, replay it."""
    eng = sb.get_engine(0)
    eng.set_sync(False)
    fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()  # binds the engine to this stream, builds the plan
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(outer):
            g.replay()
        e1.record(s)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * outer)


def col(shape):
    st, acc = [], 1
    for x in shape:
        st.append(acc)
        acc *= x
    return tuple(st)


def main():
    totals = [int(np.ceil(2 ** e)) for e in np.arange(8, 27.1, 1.5)]
    print("# permutedims!(B, A, p), Float64, s^4 elements; bytes = 2*8*s^4; device time via CUDA-graph replay")
    print(f"{'s':>5} {'elements':>11} " + " ".join(f"{'p=' + str(p):>34}" for p in ((3, 2, 1, 0), (1, 2, 3, 0), (2, 3, 0, 1))))
    for tot in totals:
        s = max(2, int(round(tot ** 0.25)))
        shape = (s,) * 4
        n = s ** 4
        a = torch.randn(n, dtype=torch.float64, device="cuda")
        b = torch.empty_like(a)
        A, B = sb.StridedView(a, shape, col(shape)), sb.StridedView(b, shape, col(shape))
        cells = []
        for p in ((3, 2, 1, 0), (1, 2, 3, 0), (2, 3, 0, 1)):
            Ap = A.permutedims(p)
            us = graph_time(lambda: sb.copy_(B, Ap))
            assert torch.equal(b.view(*shape[::-1]).permute(3, 2, 1, 0), a.view(*shape[::-1]).permute(3, 2, 1, 0).permute(*p))
            gbs = 2 * 8 * n / us / 1e3
            cells.append(f"{us:9.2f}us {gbs:8.1f}GB/s {gbs / PEAK:5.2f} {'tma' if sb.plan_describe(sb.make_desc([], 0, 0, 0.0, shape, [B, Ap])).get('tma') else 'gen'}")
        print(f"{s:>5} {n:>11} " + " ".join(f"{c:>34}" for c in cells))
    print()
    print("# sum(A), Float64 1-D (benchmark_sum); bytes = 8*n")
    for tot in [int(np.ceil(2 ** e)) for e in np.arange(8, 29.1, 1.5)]:
        a = torch.randn(tot, dtype=torch.float64, device="cuda")
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
        A = sb.StridedView(a)
        O = sb.StridedView(out, (tot,), (0,))
        us = graph_time(lambda: sb.run_mapreduce([], 1, 1, 0.0, (tot,), [O, A]))
        torch.cuda.synchronize()
        ref = a.sum().item()
        assert abs(out.item() - ref) <= 1e-9 * max(1.0, abs(ref)) * 1e3, (out.item(), ref)
        gbs = 8 * tot / us / 1e3
        print(f"{tot:>11} {us:9.2f}us {gbs:8.1f}GB/s {gbs / PEAK:5.2f}")


if __name__ == "__main__":
    main()
