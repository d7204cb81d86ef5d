"""Graph-timed probe of everyday shapes OUTSIDE the BASELINE list (tools/, not product): broadcasts, axpy, 3-D permutes,
stepped views, complex transposes.  Effective GB/s = compulsory bytes / time."""
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
K_ = lambda v: (1, 0, float(v), 0.0)  # noqa: E731
C_ = lambda f: (2, sb.abi.FN[f], 0.0, 0.0)  # noqa: E731


def col(shape):
    st, acc = [], 1
    for s in shape:
        st.append(acc)
        acc *= s
    return tuple(st)


def run(name, prog, dims, views, nbytes, reps=30):
    ms = _time(lambda i: sb.run_mapreduce(prog, 0, 0, 0.0, dims, views), reps)
    print(f"{name:52s} {ms*1e3:9.2f} us {nbytes/ms/1e6:8.1f} GB/s {nbytes/ms/1e6/PEAK:.3f}  {_kernel(prog, 0, 0, dims, views)[:70]}", flush=True)


n = 4096
a = torch.randn(n * n, dtype=torch.float64, device=dev)
b = torch.zeros(n * n, dtype=torch.float64, device=dev)
v = torch.randn(n, dtype=torch.float64, device=dev)
Bv, Av = sb.StridedView(b, (n, n), (1, n)), sb.StridedView(a, (n, n), (1, n))
run("B .= A .+ v   (v along dim 1, broadcast over columns)", [A_(0), A_(1), C_("add")], (n, n), [Bv, Av, sb.StridedView(v, (n, n), (1, 0))], 2 * n * n * 8)
run("B .= A .* v'  (v along dim 2, broadcast over rows)", [A_(0), A_(1), C_("mul")], (n, n), [Bv, Av, sb.StridedView(v, (n, n), (0, 1))], 2 * n * n * 8)
run("B .= 2 .* A .+ B  (axpy, dense 4096^2)", [K_(2), A_(0), C_("mul"), A_(1), C_("add")], (n, n), [Bv, Av, Bv], 3 * n * n * 8)
run("B .= A .* exp.(-2 .* A) .+ sin.(A .* A)  (README)", [A_(0), K_(-2), A_(1), C_("mul"), C_("exp"), C_("mul"), A_(2), A_(3), C_("mul"), C_("sin"), C_("add")], (n, n), [Bv, Av, Av, Av, Av], 2 * n * n * 8, reps=10)
run("B[1:2:end, :] .= A[1:2:end, :]  (stepped rows)", [], (n // 2, n), [sb.StridedView(b, (n // 2, n), (2, n)), sb.StridedView(a, (n // 2, n), (2, n))], 2 * (n // 2) * n * 8)
run("B[:, 1:2:end] .= A[:, 1:2:end]  (stepped columns)", [], (n, n // 2), [sb.StridedView(b, (n, n // 2), (1, 2 * n)), sb.StridedView(a, (n, n // 2), (1, 2 * n))], 2 * n * (n // 2) * 8)
run("B .= reverse(A, dims=1)  (negative stride)", [], (n, n), [Bv, sb.StridedView(a, (n, n), (-1, n), offset=n - 1)], 2 * n * n * 8)
m = 256
sh = (m, m, m)
a3 = torch.randn(m ** 3, dtype=torch.float64, device=dev)
b3 = torch.zeros(m ** 3, dtype=torch.float64, device=dev)
for p in ((1, 0, 2), (2, 1, 0), (0, 2, 1), (1, 2, 0), (2, 0, 1)):
    run(f"permutedims!(B, A, {tuple(q + 1 for q in p)})  256^3 f64", [], sh, [sb.StridedView(b3, sh, col(sh)), sb.StridedView(a3, sh, col(sh)).permutedims(p)], 2 * m ** 3 * 8)
ac = torch.view_as_complex(torch.randn(n * n, 2, dtype=torch.float32, device=dev))
bc = torch.zeros_like(ac)
run("adjoint!(B, A)  4096^2 ComplexF32", [], (n, n), [sb.StridedView(bc, (n, n), (1, n)), sb.StridedView(ac, (n, n), (n, 1), conj=True)], 2 * n * n * 8)
az = torch.view_as_complex(torch.randn(n * n // 4, 2, dtype=torch.float64, device=dev))
bz = torch.zeros_like(az)
h = n // 2
run("transpose!(B, A)  2048^2 ComplexF64", [], (h, h), [sb.StridedView(bz, (h, h), (1, h)), sb.StridedView(az, (h, h), (h, 1))], 2 * h * h * 16)
af = torch.randn(n * n, dtype=torch.float32, device=dev)
bd = torch.zeros(n * n, dtype=torch.float64, device=dev)
run("B(Float64) .= A(Float32)'  (mixed eltypes)", [], (n, n), [sb.StridedView(bd, (n, n), (1, n)), sb.StridedView(af, (n, n), (n, 1))], n * n * 12)
