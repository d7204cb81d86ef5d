"""Graph-timed probe of reductions outside the BASELINE list (tools/, not product)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import strided_jl_b200 as sb  # noqa: E402
from bench_configs import _time, _kernel  # noqa: E402

PEAK = 6545.9
dev = torch.device("cuda", 0)
eng = sb.get_engine(0)
eng.set_sync(False)
A_ = lambda i: (0, i, 0.0, 0.0)  # noqa: E731
C_ = lambda f: (2, sb.abi.FN[f], 0.0, 0.0)  # noqa: E731


def col(shape):
    st, acc = [], 1
    for s in shape:
        st.append(acc)
        acc *= s
    return tuple(st)


def run(name, prog, op, dims, views, nbytes, want=None, reps=30):
    ms = _time(lambda i: sb.run_mapreduce(prog, op, 1, 0.0, dims, views), reps)
    err = ""
    if want is not None:
        got = views[0].parent
        err = f"relerr {float((got - want).abs().max() / want.abs().max()):.1e}"
    print(f"{name:52s} {ms*1e3:9.2f} us {nbytes/ms/1e6:8.1f} GB/s {nbytes/ms/1e6/PEAK:.3f} {err}  {_kernel(prog, op, 1, dims, views)[:70]}", flush=True)


m = 64
sh = (m,) * 4
a = torch.randn(m ** 4, dtype=torch.float64, device=dev)
o1 = torch.zeros(1, dtype=torch.float64, device=dev)
Z = (0, 0, 0, 0)
run("sum(A) dense 64^4", [], 1, sh, [sb.StridedView(o1, sh, Z), sb.StridedView(a, sh, col(sh))], m ** 4 * 8, a.sum().view(1))
run("sum(permutedims(A,(4,2,1,3))) 64^4", [], 1, sh, [sb.StridedView(o1, sh, Z), sb.StridedView(a, sh, col(sh)).permutedims((3, 1, 0, 2))], m ** 4 * 8, a.sum().view(1))
run("maximum(abs, A) 64^4", [A_(0), C_("abs")], 4, sh, [sb.StridedView(o1, sh, Z), sb.StridedView(a, sh, col(sh))], m ** 4 * 8, a.abs().max().view(1))
n = 1 << 25
x, y = torch.randn(n, dtype=torch.float64, device=dev), torch.randn(n, dtype=torch.float64, device=dev)
run("sum(x .* y) 2^25 (dot)", [A_(0), A_(1), C_("mul")], 1, (n,), [sb.StridedView(o1, (n,), (0,)), sb.StridedView(x), sb.StridedView(y)], 2 * n * 8, (x * y).sum().view(1))
run("sum(x[1:2:end]) 2^24 of 2^25", [], 1, (n // 2,), [sb.StridedView(o1, (n // 2,), (0,)), sb.StridedView(x, (n // 2,), (2,))], n // 2 * 8, x[::2].sum().view(1))
o2 = torch.zeros(m * m, dtype=torch.float64, device=dev)
A4 = a.view(m, m, m, m)  # torch index order is reversed: A4[l,k,j,i] = A[i,j,k,l]
run("sum(A; dims=(1,2)) 64^4 -> 64x64", [], 1, sh, [sb.StridedView(o2, sh, (0, 0, 1, m)), sb.StridedView(a, sh, col(sh))], m ** 4 * 8, A4.sum(dim=(2, 3)).contiguous().view(-1))
run("sum(A; dims=(3,4)) 64^4 -> 64x64", [], 1, sh, [sb.StridedView(o2, sh, (1, m, 0, 0)), sb.StridedView(a, sh, col(sh))], m ** 4 * 8, A4.sum(dim=(0, 1)).contiguous().view(-1))
run("sum(A; dims=(1,3)) 64^4 -> 64x64", [], 1, sh, [sb.StridedView(o2, sh, (0, 1, 0, m)), sb.StridedView(a, sh, col(sh))], m ** 4 * 8, A4.sum(dim=(1, 3)).contiguous().view(-1))
run("sum(A; dims=(2,4)) 64^4 -> 64x64", [], 1, sh, [sb.StridedView(o2, sh, (1, 0, m, 0)), sb.StridedView(a, sh, col(sh))], m ** 4 * 8, A4.sum(dim=(0, 2)).contiguous().view(-1))
o3 = torch.zeros(m ** 3, dtype=torch.float64, device=dev)
run("sum(A; dims=4) 64^4 -> 64^3", [], 1, sh, [sb.StridedView(o3, sh, (1, m, m * m, 0)), sb.StridedView(a, sh, col(sh))], m ** 4 * 8, A4.sum(dim=0).contiguous().view(-1))
run("sum(A; dims=1) 64^4 -> 64^3", [], 1, sh, [sb.StridedView(o3, sh, (0, 1, m, m * m)), sb.StridedView(a, sh, col(sh))], m ** 4 * 8, A4.sum(dim=3).contiguous().view(-1))
