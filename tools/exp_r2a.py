"""Round-2 experiment batch A (one gpurun call):
  1. plan variants for the launch-bound configs (C3 / C1 through the TMA ring, C4' variants), graph-timed;
  2. ZERO-COPY end to end: the kernels read pinned HOST memory directly (UVA pointers passed to sb_mapreduce) and write
     the result straight into pinned host memory -- H2D and D2H overlap inside ONE kernel on the full-duplex link --
     against the staged path of sb_mapreduce_host (H2D copy, kernel, D2H copy).

    python tools/exp_r2a.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import strided_jl_b200 as sb  # noqa: E402
from strided_jl_b200.engine import Engine  # noqa: E402
from tools.exp_orbit import time_variant  # noqa: E402
from tools.profile_case import MAKE  # noqa: E402

VARIANTS = {
    "c3": [{}, {"SB_FORCE_EPT": "8"}, {"SB_FORCE_EPT": "8", "SB_TMA_STAGES": "2"}, {"SB_FORCE_EPT": "16"}],
    "c1": [{}, {"SB_FORCE_EPT": "8"}, {"SB_FORCE_EPT": "8", "SB_TMA_STAGES": "2"}],
    "c4p": [{}, {"SB_ORBIT_STAGING": "3"}, {"SB_NO_ORBIT": "1"}, {"SB_NO_ORBIT": "1", "SB_FORCE_EPT": "8"}],
    "c4": [{}, {"SB_ORBIT_DIRECT": "2"}, {"SB_ORBIT_DIRECT": "1"}],
    "c5shard": [{}],
}


def clear_env():
    for k in list(os.environ):
        if k.startswith("SB_"):
            del os.environ[k]


def zero_copy():
    n = 4000
    c = MAKE["c2"]()
    a_host = torch.from_numpy(c.parents[1]).pin_memory()
    b_host = torch.zeros(n * n, dtype=torch.float64).pin_memory()
    want = ((c.parents[1].reshape(n, n) + c.parents[1].reshape(n, n).T) / 2).reshape(-1)  # column-major flat == row-major of the transpose; symmetric anyway
    alg = 2 * n * n * 8
    for env in ({}, {"SB_ORBIT_NIN2": "1"}, {"SB_NO_TMA": "1"}, {"SB_ORBIT_NIN2": "1", "SB_ORBIT_DIRECT": "1"}):
        clear_env()
        os.environ.update(env)
        eng = Engine(0)
        eng.set_sync(True)
        try:
            A = sb.StridedView(a_host, (n, n), (1, n))
            B = sb.StridedView(b_host, (n, n), (1, n))
            desc = sb.make_desc(c.tokens, 0, 0, 0.0, (n, n), [B, A, A.T])
            b_host.zero_()
            eng.mapreduce(desc, host=False)  # UVA: the pinned host pointers ARE device-accessible addresses
            ok = np.array_equal(b_host.numpy(), want)
            ts = []
            for _ in range(5):
                t0 = time.perf_counter()
                eng.mapreduce(desc, host=False)
                ts.append(time.perf_counter() - t0)
            t = min(ts)
            print(f"zero-copy c2 env={env} ok={ok} one sync call {t * 1e3:.3f} ms = {alg / t / 1e9:.1f} GB/s (mean {np.mean(ts) * 1e3:.3f} ms)", flush=True)
        except Exception as e:
            print(f"zero-copy c2 env={env} ERROR {e}", flush=True)
        eng.close()
    # staged path (what bench.py's e2e measures today), one synchronous call at a time
    clear_env()
    eng = Engine(0)
    eng.set_sync(True)
    Ah = sb.StridedView(a_host.numpy(), (n, n), (1, n))
    Bh = sb.StridedView(b_host.numpy(), (n, n), (1, n))
    for _ in range(2):
        sb.run_mapreduce(c.tokens, 0, 0, 0.0, (n, n), [Bh, Ah, Ah.T], engine=eng)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        sb.run_mapreduce(c.tokens, 0, 0, 0.0, (n, n), [Bh, Ah, Ah.T], engine=eng)
        ts.append(time.perf_counter() - t0)
    print(f"staged sb_mapreduce_host one sync call {min(ts) * 1e3:.3f} ms = {alg / min(ts) / 1e9:.1f} GB/s", flush=True)
    # raw link: cudaMemcpyAsync H2D and D2H of 128 MB, alone and concurrently
    d = torch.empty(n * n, dtype=torch.float64, device="cuda")
    d2 = torch.empty(n * n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for name, fn in (("h2d", lambda: d.copy_(a_host, non_blocking=True)), ("d2h", lambda: b_host.copy_(d2, non_blocking=True))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) / 5
        print(f"raw {name} 128 MB: {t * 1e3:.3f} ms = {n * n * 8 / t / 1e9:.1f} GB/s", flush=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s1):
            d.copy_(a_host, non_blocking=True)
        with torch.cuda.stream(s2):
            b_host.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / 5
    print(f"raw h2d + d2h concurrently, 128 MB each way: {t * 1e3:.3f} ms = {alg / t / 1e9:.1f} GB/s (both directions)", flush=True)
    eng.close()


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
                orb = p.get("orbit")
                print(f"{nm} env={env} us={us:.2f} same_as_default={same} family={p.get('family')} ept={p.get('ept')} orbit={ {k: orb[k] for k in ('items', 'nstage', 'nstaging', 'tile_bytes')} if orb else 0} tile={p.get('tile')} tma={p.get('tma')}", flush=True)
            except Exception as e:
                print(f"{nm} env={env} ERROR {e}", flush=True)
        del dev
    zero_copy()


if __name__ == "__main__":
    main()
