"""Round-2 experiment batch J: TMA ring vs LSU kernel on odd extents (128-byte swizzled rows), shifted last tile on/off."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from helpers import case_c2, case_c3  # noqa: E402
from tools.exp_orbit import time_variant  # noqa: E402

CASES = {"rev41": lambda: case_c3(41), "rev54": lambda: case_c3(54), "rev70": lambda: case_c3(70), "rev91": lambda: case_c3(91),
         "rot70": lambda: case_c3(70, p=(1, 2, 3, 0)), "c2_4002": lambda: case_c2(4002), "c2_3001": lambda: case_c2(3001)}
VARS = [{}, {"SB_NO_TMA": "1"}, {"SB_NO_SHIFT": "1"}, {"SB_NO_TMA": "1", "SB_NO_SHIFT": "1"}]


def main():
    for nm, mk in CASES.items():
        c = mk()
        dev = [torch.from_numpy(p).cuda() for p in c.parents]
        first = None
        nbytes = sum(p.nbytes for p in c.parents)
        for env in VARS:
            try:
                dev[c.views[0].parent].zero_()
                us, p = time_variant(c, dev, env, 20)
                got = dev[c.views[0].parent].cpu().numpy()
                if first is None:
                    first = got
                print(f"{nm} env={env} us={us:.2f} GB/s={nbytes / us * 1e-3:.0f} same={bool(np.array_equal(got, first))} tile={p.get('tile')} tma={p.get('tma')} shift={p.get('shift_last')}", flush=True)
            except Exception as e:
                print(f"{nm} env={env} ERROR {e}", flush=True)
        del dev


if __name__ == "__main__":
    main()
