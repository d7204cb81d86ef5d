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

N2 = {"SB_ORBIT_NIN2": "1"}
VARIANTS = {
    "c2": [{}, {"SB_NO_PDL": "1"}, dict(N2)],
    "c4": [{}, {"SB_NO_PDL": "1"}],
    "c4p": [{}, {"SB_NO_PDL": "1"}],
    "c1": [{}, {"SB_NO_PDL": "1"}],
    "c3": [{}, {"SB_NO_PDL": "1"}],
    "c5": [{}, {"SB_NO_PDL": "1"}],
    "c5shard": [{}, {"SB_NO_PDL": "1"}],
}


def time_variant(c, dev, env, reps):
    """us per call, device time: a CUDA graph of `reps` launches replayed (no host launch overhead in the number)"""
    for k in list(os.environ):
        if k.startswith("SB_"):
            del os.environ[k]
    os.environ.update(env)
    eng = Engine(0)
    eng.set_sync(False)
    views = c._svs(dev)
    st = torch.cuda.Stream()
    best = 1e30
    with torch.cuda.stream(st):
        eng.set_stream(st.cuda_stream)
        for _ in range(3):
            sb.run_mapreduce(c.tokens, c.op, c.initop, c.init, c.dims, views, engine=eng)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                sb.run_mapreduce(c.tokens, c.op, c.initop, c.init, c.dims, views, engine=eng)
        g.replay()
        torch.cuda.synchronize()
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(5):
                g.replay()
            e1.record(st)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / (5 * reps))
    p = c.plan()
    del g
    eng.close()
    return best, p


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
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
