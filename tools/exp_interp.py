"""Interpreter vs recipe: same memory traffic, different element-function path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import strided_jl_b200 as sb

def t(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

sb.get_engine(0).set_sync(False)
for dt, esz in ((torch.float64, 8), (torch.float32, 4)):
    n = 4096
    x, y, z, o = (torch.randn(n * n, dtype=dt, device="cuda") for _ in range(4))
    X, Y, Z, O = (sb.StridedView(v, (n, n), (1, n)) for v in (x, y, z, o))
    cases = {
        "recipe  (x+y)+z        dense": lambda: O.assign((X + Y) + Z),
        "interp  x+(y+z)        dense": lambda: O.assign(X + (Y + Z)),
        "interp  x*y-z          dense": lambda: O.assign(X * Y - Z),
        "interp  sin(x)+y/exp(-|z|)  ": lambda: O.assign(sb.sin(X) + Y / sb.exp(-sb.abs_(Z))),
        "recipe  x+y'           transp": lambda: O.assign(X + Y.T),
        "interp  x-y'           transp": lambda: O.assign(X - Y.T),
        "interp  2x-y'          transp": lambda: O.assign(2 * X - Y.T),
    }
    for name, fn in cases.items():
        us = t(fn)
        nin = 3 if "z" in name else 2
        print(f"{dt} {name}: {us:8.1f} us  {(nin + 1) * n * n * esz / us / 1e3:7.0f} GB/s (operand bytes)")
