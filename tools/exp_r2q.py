"""Round-2 experiment batch Q: balanced tiles (MapParams::umask) vs power-of-two tiles with a shifted last tile on odd extents."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from helpers import case_c1, case_c2, case_c3  # noqa: E402
from tools.exp_orbit import time_variant  # noqa: E402

CASES = {"rev41": lambda: case_c3(41), "rev54": lambda: case_c3(54), "rev70": lambda: case_c3(70), "rev91": lambda: case_c3(91), "rev100": lambda: case_c3(100),
         "rot70": lambda: case_c3(70, p=(1, 2, 3, 0)), "swap91": lambda: case_c3(91, p=(2, 3, 0, 1)), "c2_3001": lambda: case_c2(3001), "c1_1001": lambda: case_c1(1001), "c3_32": lambda: case_c3(32), "rev64": lambda: case_c3(64), "rot64": lambda: case_c3(64, p=(1, 2, 3, 0)),
         "rev128": lambda: case_c3(128), "c1_1000": lambda: case_c1(1000), "c2_4000": lambda: case_c2(4000), "c2_4002": lambda: case_c2(4002), "c2_4096": lambda: case_c2(4096), "c2_2048": lambda: case_c2(2048), "c1_1024": lambda: case_c1(1024), "c1_4096": lambda: case_c1(4096), "rev48": lambda: case_c3(48), "rev96": lambda: case_c3(96), "c1_3000": lambda: case_c1(3000), "f32_rev32": lambda: case_c3(32, np.float32), "f32_rev64": lambda: case_c3(64, np.float32), "f32_rev128": lambda: case_c3(128, np.float32), "f32_c1_1000": lambda: case_c1(1000, np.float32), "f32_c1_4096": lambda: case_c1(4096, np.float32), "f32_rot64": lambda: case_c3(64, np.float32, p=(1, 2, 3, 0)), "f32_c1_8192": lambda: case_c1(8192, np.float32), "f32_rev70": lambda: case_c3(70, np.float32), "f32_rev54": lambda: case_c3(54, np.float32), "f32_rev100": lambda: case_c3(100, np.float32), "f32_c1_3000": lambda: case_c1(3000, np.float32), "f32_c2_4000": lambda: case_c2(4000, np.float32), "f32_c2_4096": lambda: case_c2(4096, np.float32), "f32_c2_1000": lambda: case_c2(1000, np.float32), "f32_rev48": lambda: case_c3(48, np.float32), "f32_c1_2000": lambda: case_c1(2000, np.float32), "f32_rev40": lambda: case_c3(40, np.float32), "rot91": lambda: case_c3(91, p=(1, 2, 3, 0)), "rot100": lambda: case_c3(100, p=(1, 2, 3, 0)), "lrot70": lambda: case_c3(70, p=(3, 0, 1, 2)), "lrot91": lambda: case_c3(91, p=(3, 0, 1, 2)), "t3001": lambda: case_c1(3001), "t5001": lambda: case_c1(5001), "f32_rev41": lambda: case_c3(41, np.float32), "f32_rev91": lambda: case_c3(91, np.float32), "f32_t3001": lambda: case_c1(3001, np.float32), "rev55": lambda: case_c3(55), "rev27": lambda: case_c3(27), "rev59": lambda: case_c3(59), "c1_119": lambda: case_c1(119), "f32_c1_955": lambda: case_c1(955, np.float32)}
VARS = [{}, {"SB_NO_BALANCED": "1"}]


def main():
    only = sys.argv[1].split(",") if len(sys.argv) > 1 else list(CASES)
    for nm in only:
        c = CASES[nm]()
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
                print(f"{nm} env={env} us={us:.2f} GB/s={nbytes / us * 1e-3:.0f} frac={nbytes / us * 1e-3 / 6545.9:.3f} same={bool(np.array_equal(got, first))} tile={p.get('tile')} "
                      f"bal={p.get('balanced')} tma={p.get('tma')} lsu_desc={p.get('lsu_desc')} order={p.get('tile_order')} shift={p.get('shift_last')} ept={p.get('ept')}", flush=True)
            except Exception as e:
                print(f"{nm} env={env} ERROR {e}", flush=True)
        del dev


if __name__ == "__main__":
    main()
