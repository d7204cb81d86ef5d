"""Run one BASELINE config a few times on cuda:0 (for ncu / compute-sanitizer captures under gpurun).

    python tools/profile_case.py c2|c1|c3|c4|c4p|c5|c5shard|all [reps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import strided_jl_b200 as sb  # noqa: E402
from helpers import case_c1, case_c2, case_c3, case_c4, case_c5  # noqa: E402

MAKE = {
    "c1": lambda: case_c1(1000), "c2": lambda: case_c2(4000), "c3": lambda: case_c3(32), "c4": lambda: case_c4(64),
    "c4p": lambda: case_c4(32, np.float64), "c4s": lambda: case_c4(16), "c4e": lambda: case_c4(20), "c2s": lambda: case_c2(512), "c5": lambda: case_c5(8, 4096), "c5shard": lambda: case_c5(1, 4096),
    "rev91": lambda: case_c3(91), "rev70": lambda: case_c3(70), "c5small": lambda: case_c5(8, 512),
}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c2"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    names = list(MAKE) if which == "all" else [which]
    for nm in names:
        c = MAKE[nm]()
        dev = [torch.from_numpy(p).cuda() for p in c.parents]
        views = c._svs(dev)
        for _ in range(reps):
            sb.run_mapreduce(c.tokens, c.op, c.initop, c.init, c.dims, views)
        torch.cuda.synchronize()
        print(nm, "done", sb.get_engine(0).stats())


if __name__ == "__main__":
    main()
