"""A/B of two builds of libstrided_b200.so on the SAME box: times config 2 (and config 3 / config 4) through a given
shared library, graph-replayed.  Only the C ABI is used (the struct layout of sb_desc is the same in all builds).

    python tools/ab_lib.py <path/to/libstrided_b200.so> [reps]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from strided_jl_b200 import abi  # noqa: E402  (struct definitions only)
from strided_jl_b200.view import StridedView  # noqa: E402
from strided_jl_b200.engine import make_desc  # noqa: E402


def main():
    path = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.sb_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.sb_ctx_set_stream.argtypes = [vp, vp]
    lib.sb_ctx_set_sync.argtypes = [vp, C.c_int]
    lib.sb_mapreduce.argtypes = [vp, C.POINTER(abi.sb_desc)]
    ctx = vp()
    assert lib.sb_ctx_create(0, None, C.byref(ctx)) == 0
    lib.sb_ctx_set_sync(ctx, 0)
    dev = torch.device("cuda", 0)
    A_ = lambda i: (0, i, 0.0, 0.0)  # noqa: E731
    CALL = lambda f: (2, abi.FN[f], 0.0, 0.0)  # noqa: E731

    def col(shape):
        st, acc = [], 1
        for s in shape:
            st.append(acc)
            acc *= s
        return tuple(st)

    jobs = []
    n = 4000
    a = torch.randn(n * n, dtype=torch.float64, device=dev)
    b = torch.empty_like(a)
    Av, Bv = StridedView(a, (n, n), (1, n)), StridedView(b, (n, n), (1, n))
    jobs.append(("c2", make_desc([A_(0), A_(1), CALL("add"), (1, 0, 2.0, 0.0), CALL("div")], 0, 0, 0.0, (n, n), [Bv, Av, Av.T]), 2 * n * n * 8))
    m = 32
    a3 = torch.randn(m ** 4, dtype=torch.float64, device=dev)
    b3 = torch.empty_like(a3)
    sh = (m,) * 4
    jobs.append(("c3", make_desc([], 0, 0, 0.0, sh, [StridedView(b3, sh, col(sh)), StridedView(a3, sh, col(sh)).permutedims((3, 2, 1, 0))]), 2 * m ** 4 * 8))
    m = 64
    a4 = torch.randn(m ** 4, dtype=torch.float32, device=dev)
    b4 = torch.empty_like(a4)
    sh4 = (m,) * 4
    A4 = StridedView(a4, sh4, col(sh4))
    jobs.append(("c4", make_desc([A_(0), A_(1), CALL("add"), A_(2), CALL("add"), A_(3), CALL("add")], 0, 0, 0.0, sh4,
                                 [StridedView(b4, sh4, col(sh4))] + [A4.permutedims(p) for p in ((0, 1, 2, 3), (1, 2, 3, 0), (2, 3, 0, 1), (3, 0, 1, 2))]), 2 * m ** 4 * 4))
    m = 1000
    a1 = torch.randn(m * m, dtype=torch.float64, device=dev)
    b1 = torch.empty_like(a1)
    jobs.append(("c1", make_desc([(1, 0, 3.0, 0.0), A_(0), CALL("mul")], 0, 0, 0.0, (m, m), [StridedView(b1, (m, m), (1, m)), StridedView(a1, (m, m), (m, 1))]), 2 * m * m * 8))
    keep = []
    for m in (41, 70, 91):  # odd extents, reversal permutation (benchmarks/benchtests.jl:40)
        ao = torch.randn(m ** 4, dtype=torch.float64, device=dev)
        bo = torch.empty_like(ao)
        keep.append((ao, bo))
        sho = (m,) * 4
        jobs.append((f"rev{m}", make_desc([], 0, 0, 0.0, sho, [StridedView(bo, sho, col(sho)), StridedView(ao, sho, col(sho)).permutedims((3, 2, 1, 0))]), 2 * m ** 4 * 8))
    st = torch.cuda.Stream()
    for name, desc, nbytes in jobs:
        with torch.cuda.stream(st):
            lib.sb_ctx_set_stream(ctx, vp(st.cuda_stream))
            for _ in range(3):
                assert lib.sb_mapreduce(ctx, C.byref(desc)) == 0
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for _ in range(reps):
                    assert lib.sb_mapreduce(ctx, C.byref(desc)) == 0
            g.replay()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                g.replay()
                e1.record(st)
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
            del g
        print(f"{os.path.basename(path)} {name}: {best:.2f} us  {nbytes / best * 1e-3:.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
