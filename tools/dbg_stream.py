"""Debug aid: run a few streamed-reduction cases on the GPU and print got / want (tools/, not product)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import cases  # noqa: E402
from helpers import case_c5  # noqa: E402

want_names = sys.argv[1:] or ["sum_100_100_2_float64", "stream_dims12_float64", "stream_inter8_float64", "stream_inter2_float64"]
allc = {c.name: c for c in cases.all_cases(1.0)}
allc["c5_8_256"] = case_c5(8, 256)
for nm in want_names:
    c = allc[nm]
    got = c.run_gpu("device")
    want = c.expected()
    p = c.plan()
    print(nm, "env", {k: v for k, v in os.environ.items() if k.startswith("SB_")}, "stream", p.get("stream"))
    print("  got ", np.array2string(got.ravel()[:8], precision=6))
    print("  want", np.array2string(want.ravel()[:8], precision=6))
