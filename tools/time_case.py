"""Event-timed device-resident timing of one BASELINE config (tuning aid; honours SB_FORCE_EPT etc.).

    python tools/time_case.py c2 [reps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import strided_jl_b200 as sb  # noqa: E402
from tools.profile_case import MAKE  # noqa: E402


def main():
    which = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    c = MAKE[which]()
    dev = [torch.from_numpy(p).cuda() for p in c.parents]
    views = c._svs(dev)
    eng = sb.get_engine(0)
    eng.set_sync(False)
    for _ in range(10):
        sb.run_mapreduce(c.tokens, c.op, c.initop, c.init, c.dims, views)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sb.run_mapreduce(c.tokens, c.op, c.initop, c.init, c.dims, views)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    p = c.plan()
    print(f"{which} env={ {k: v for k, v in os.environ.items() if k.startswith('SB_')} } us={us:.2f} tile={p.get('tile')} ept={p.get('ept')} grid={p.get('grid')}")


if __name__ == "__main__":
    main()
