"""Round-2 experiment batch P: orbit split factor on the one-wave config C4' (F64 32^4 4-way sum)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from tools.exp_orbit import time_variant  # noqa: E402
from tools.profile_case import MAKE  # noqa: E402

VARIANTS = {"c4p": [{}, {"SB_ORBIT_SPLIT": "1"}, {"SB_ORBIT_SPLIT": "2"}, {"SB_ORBIT_SPLIT": "4"}, {"SB_NO_ORBIT": "1"}]}
for nm, envs in VARIANTS.items():
    c = MAKE[nm]()
    dev = [torch.from_numpy(p).cuda() for p in c.parents]
    for env in envs:
        try:
            us, p = time_variant(c, dev, env, 50)
            print(f"{nm} env={env} us={us:.2f} orbit={p.get('orbit') or 0} tile={p.get('tile')} tma={p.get('tma')}", flush=True)
        except Exception as e:
            print(f"{nm} env={env} ERROR {e}", flush=True)
