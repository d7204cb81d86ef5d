"""Dense-copy bandwidth of the generic kernel as a function of the per-lane access width (4/8/16 B)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import strided_jl_b200 as sb

def t(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

eng = sb.get_engine(0); eng.set_sync(False)
nbytes = 1 << 29
for dt, esz in ((torch.float32, 4), (torch.float64, 8), (torch.complex64, 8), (torch.complex128, 16)):
    n = nbytes // esz
    x = torch.zeros(n, dtype=dt, device="cuda"); y = torch.empty_like(x)
    X, Y = sb.StridedView(x), sb.StridedView(y)
    ms = t(lambda: sb.copy_(Y, X))
    print(f"engine copy {dt} {2*nbytes/ms/1e6:.0f} GB/s   ({ms*1e3:.1f} us)")
    ms = t(lambda: y.copy_(x))
    print(f"torch  copy {dt} {2*nbytes/ms/1e6:.0f} GB/s")
    # 2-D transpose copy
    m = int((n) ** 0.5) // 64 * 64
    Xm = sb.StridedView(x, (m, m), (1, m)); Ym = sb.StridedView(y, (m, m), (1, m))
    ms = t(lambda: sb.copy_(Ym, Xm.T))
    print(f"engine transpose {dt} {m}x{m} {2*m*m*esz/ms/1e6:.0f} GB/s  plan={sb.plan_describe(sb.make_desc([],0,0,0.0,(m,m),[Ym,Xm.T])).get('tma')}")
    del x, y
