"""Round-2 experiment batch E: A/B of the tile-record prefetch in the TMA ring kernel, stream-reduction variants."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from tools.exp_orbit import time_variant  # noqa: E402
from tools.profile_case import MAKE  # noqa: E402

NP = {"SB_TMA_NOPREFETCH": "1"}
VARIANTS = {
    "c2": [{}, dict(NP), {}, dict(NP)],
    "c1": [{}, dict(NP), {}, dict(NP)],
    "c3": [{}, dict(NP), {}, dict(NP)],
    "c5shard": [{}, {"SB_NO_STREAM": "1"}, {"SB_STREAM_STAGES": "6"}, {"SB_STREAM_CHUNK": "16384", "SB_STREAM_STAGES": "8"}, {}],
    "c5": [{}],
}


def main():
    for nm, vs in VARIANTS.items():
        c = MAKE[nm]()
        dev = [torch.from_numpy(p).cuda() for p in c.parents]
        first = None
        for env in vs:
            try:
                dev[c.views[0].parent].zero_()
                us, p = time_variant(c, dev, env, 40)
                got = dev[c.views[0].parent].cpu().numpy()
                if first is None:
                    first = got
                same = bool(np.array_equal(got, first)) if c.op == 0 else bool(np.allclose(got, first, rtol=1e-9))
                print(f"{nm} env={env} us={us:.2f} same_as_default={same} tma={p.get('tma')} stream={p.get('stream')}", flush=True)
            except Exception as e:
                print(f"{nm} env={env} ERROR {e}", flush=True)
        del dev


if __name__ == "__main__":
    main()
