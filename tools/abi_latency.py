"""Host-side cost of one sb_mapreduce call (plan-cache hit, stream-ordered), measured through ctypes with a prebuilt
descriptor: what a Julia `ccall` would pay per `@strided` statement on top of the kernel."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import strided_jl_b200 as sb
from helpers import case_c2, case_c3, case_c5


def main():
    eng = sb.get_engine(0)
    eng.set_sync(False)
    lib = eng.lib
    for mk, name in ((lambda: case_c3(8), "tiny permute 8^4 (generic kernel)"), (lambda: case_c2(512), "C2-like 512^2 (TMA kernel: + 2 tensor-map encodes)"),
                     (lambda: case_c5(8, 64), "small reduce (2 launches)")):
        c = mk()
        dev = [torch.from_numpy(p).cuda() for p in c.parents]
        d = sb.make_desc(c.tokens, c.op, c.initop, c.init, c.dims, c._svs(dev))
        for _ in range(20):
            lib.sb_mapreduce(eng.ctx, C.byref(d))
        torch.cuda.synchronize()
        n = 3000
        t0 = time.perf_counter()
        for _ in range(n):
            lib.sb_mapreduce(eng.ctx, C.byref(d))
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"{name}: host {1e6 * (t1 - t0) / n:.2f} us/call (ctypes overhead included), drained after {1e6 * (t2 - t0) / n:.2f} us/call")


if __name__ == "__main__":
    main()
