"""Event-timed comparison of plan variants (fresh sb_ctx per variant, so plan-time env knobs take effect).

    python tools/exp_orbit.py [reps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import strided_jl_b200 as sb  # noqa: E402
from strided_jl_b200.engine import Engine  # noqa: E402
from tools.profile_case import MAKE  # noqa: E402

VARIANTS = {
    "c2": [{}, {"SB_NO_ORBIT": "1"}, {"SB_ORBIT_BITS": "6"}, {"SB_ORBIT_STAGES": "2"}, {"SB_ORBIT_STAGES": "3"}, {"SB_ORBIT_STAGES": "6"},
           {"SB_ORBIT_BITS": "6", "SB_ORBIT_STAGES": "1"}],
    "c4": [{}, {"SB_NO_ORBIT": "1"}, {"SB_ORBIT_STAGES": "1"}, {"SB_ORBIT_STAGES": "3"}],
    "c4p": [{}, {"SB_NO_ORBIT": "1"}],
    "c1": [{}],
    "c3": [{}],
}


def time_variant(c, dev, env, reps):
    for k in list(os.environ):
        if k.startswith("SB_"):
            del os.environ[k]
    os.environ.update(env)
    eng = Engine(0)
    eng.set_stream(torch.cuda.current_stream(0).cuda_stream)
    eng.set_sync(False)
    views = c._svs(dev)
    for _ in range(10):
        sb.run_mapreduce(c.tokens, c.op, c.initop, c.init, c.dims, views, engine=eng)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            sb.run_mapreduce(c.tokens, c.op, c.initop, c.init, c.dims, views, engine=eng)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    p = c.plan()
    eng.close()
    return best, p


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    which = sys.argv[2].split(",") if len(sys.argv) > 2 else list(VARIANTS)
    for nm in which:
        c = MAKE[nm]()
        dev = [torch.from_numpy(p).cuda() for p in c.parents]
        for env in VARIANTS[nm]:
            try:
                us, p = time_variant(c, dev, env, reps)
                orb = p.get("orbit")
                print(f"{nm} env={env} us={us:.2f} orbit={orb if orb else 0} tile={p.get('tile')} tma={p.get('tma')}", flush=True)
            except Exception as e:  # keep going: one bad variant must not cost the GPU call
                print(f"{nm} env={env} ERROR {e}", flush=True)
        del dev


if __name__ == "__main__":
    main()
