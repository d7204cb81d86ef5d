"""Graph-timed partial reductions of a column-major matrix / 3-D array through the generic reduce_tile kernel (tools/, not product)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import strided_jl_b200 as sb  # noqa: E402
from bench_configs import _time, _kernel  # noqa: E402

PEAK = 6545.9
dev = torch.device("cuda", 0)
eng = sb.get_engine(0)
eng.set_sync(False)
A_ = lambda i: (0, i, 0.0, 0.0)  # noqa: E731


def run(name, shape, ostrides, dt=torch.float64):
    n = int(np.prod(shape))
    a = torch.randn(n, dtype=dt, device=dev)
    nout = int(np.prod([s for s, st in zip(shape, ostrides) if st != 0]))
    o = torch.zeros(nout, dtype=dt, device=dev)
    st, acc = [], 1
    for s in shape:
        st.append(acc)
        acc *= s
    views = [sb.StridedView(o, shape, ostrides), sb.StridedView(a, shape, tuple(st))]
    ms = _time(lambda i: sb.run_mapreduce([], 1, 1, 0.0, shape, views), 50)
    want = a.view(*reversed(shape)).permute(*reversed(range(len(shape))))
    red = tuple(i for i, s in enumerate(ostrides) if s == 0)
    ref = want.sum(dim=red).contiguous().view(-1) if nout > 1 else want.sum().view(1)
    # column-major flattening of the kept dims
    kept = [s for s, st_ in zip(shape, ostrides) if st_ != 0]
    got = o.view(*reversed(kept)).permute(*reversed(range(len(kept)))).contiguous().view(-1) if len(kept) > 1 else o
    err = float((got - ref).abs().max() / ref.abs().max())
    by = n * a.element_size()
    print(f"{name:44s} {ms*1e3:8.2f} us {by/ms/1e6:8.1f} GB/s {by/ms/1e6/PEAK:.3f} relerr {err:.1e}  {_kernel([], 1, 1, shape, views)[:80]}", flush=True)


run("sum(A; dims=2)  4096x4096 (row sums)", (4096, 4096), (1, 0))
run("sum(A; dims=1)  4096x4096 (column sums)", (4096, 4096), (0, 1))
run("sum(A; dims=2)  8192x8192", (8192, 8192), (1, 0))
run("sum(A; dims=1)  8192x8192", (8192, 8192), (0, 1))
run("sum(A; dims=(1,3)) 256x256x256", (256, 256, 256), (0, 1, 0))
run("sum(A; dims=2) 256x256x256", (256, 256, 256), (1, 0, 256))
run("sum(A; dims=(2,3)) 64x512x512", (64, 512, 512), (1, 0, 0))
run("sum(A; dims=(2,3)) 128x512x512", (128, 512, 512), (1, 0, 0))
run("sum(A; dims=1) f32 8192x8192", (8192, 8192), (0, 1), torch.float32)
run("sum(A; dims=2) f32 8192x8192", (8192, 8192), (1, 0), torch.float32)
