"""Secondary measurements for bench.py --extra: the other BASELINE configs (parity-test cases, not bench lines)
timed on the device with CUDA events, each reported as algorithmic GB/s and fraction of the HBM peak.
Configs whose working set fits in the 126 MB L2 are timed twice: `warm` (same buffers back to back, L2-resident)
and `rot` (rotating over enough distinct buffer sets to exceed 2x L2: the steady-state HBM number)."""
import numpy as np
import torch

import strided_jl_b200 as sb

L2_BYTES = 126 * 1024 * 1024


def _time(fn, reps, warm=5):
    """ms per call, device time: `reps` launches captured in one CUDA graph, replayed (no host launch cost in the number;
    eager launches from Python are host-bound below ~15 us per call)."""
    reps = min(reps, 400)
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        fn(0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for i in range(reps):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            g.replay()
            e1.record(side)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
    del g
    return best


def _col(shape):
    st, acc = [], 1
    for s in shape:
        st.append(acc)
        acc *= s
    return tuple(st)


def _kernel(tokens, op, initop, dims, views):
    """kernel family the planner picks for this call (host-side introspection, sb_plan_describe)"""
    try:
        p = sb.plan_describe(sb.make_desc(tokens, op, initop, 0.0, dims, views))
    except Exception as e:  # pragma: no cover
        return f"? ({e})"
    ct = p.get("ct")
    if p.get("orbit"):
        o = p["orbit"]
        return f"map_orbit_kernel<{ct},{p.get('recipe')},EPT={o['ept']}> items {o['items']}, {o['nstage']}-stage TMA ring + TMA store"
    if p.get("stream"):
        t = p["stream"]
        return f"reduce_stream_kernel<{ct},{p.get('recipe')}> grid {t['grid']}, {t['nstage']} x {t['chunk_bytes'] // 1024} KB cp.async.bulk ring"
    if p.get("tma"):
        return f"map_tma_kernel<{ct},{p.get('recipe')},EPT={p.get('ept')}> tile {p.get('tile')}, {p['tma']}-stage TMA ring"
    if p.get("family") == "reduce_tile":
        return f"reduce_tile_kernel<{ct},{p.get('recipe')},EPT={p.get('ept')}> grid {p.get('grid')}, nsplit {p.get('nsplit')}"
    return f"{p.get('family')}_kernel<{ct},{p.get('recipe')},EPT={p.get('ept')}> tile {p.get('tile')}"


def _entry(name, alg_bytes, ms, peak, **kw):
    g = alg_bytes / (ms * 1e-3) / 1e9
    d = {"config": name, "algorithmic_bytes": alg_bytes, "ms": ms, "GBps": g, "frac_of_peak": g / peak}
    d.update(kw)
    return d


def run_all(eng, peak):
    dev = torch.device("cuda", torch.cuda.current_device())
    out = []
    eng.set_sync(False)
    A_ = lambda i: (0, i, 0.0, 0.0)  # noqa: E731
    CALL = lambda f: (2, sb.abi.FN[f], 0.0, 0.0)  # noqa: E731

    def sets_for(bytes_per_set):
        return max(2, int(2.2 * L2_BYTES // bytes_per_set) + 1)

    # plain dense copy through the engine (calibration against the driver's torch copy peak)
    n = 1 << 27
    x, y = torch.randn(n, dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.float32, device=dev)
    X, Y = sb.StridedView(x), sb.StridedView(y)
    ms = _time(lambda i: sb.copy_(Y, X), 20)
    out.append(_entry("copy f32 2^27 (dense)", 2 * n * 4, ms, peak, kernel=_kernel([], 0, 0, Y.size, [Y, X])))
    ms = _time(lambda i: y.copy_(x), 20)
    out.append(_entry("torch copy_ f32 2^27 (reference point)", 2 * n * 4, ms, peak))
    del x, y

    # C1 on the GPU: F64 1000^2  B .= 3 .* A'
    m = 1000
    k = sets_for(2 * m * m * 8)
    As = [torch.randn(m * m, dtype=torch.float64, device=dev) for _ in range(k)]
    Bs = [torch.empty(m * m, dtype=torch.float64, device=dev) for _ in range(k)]
    prog = [(1, 0, 3.0, 0.0), A_(0), CALL("mul")]
    views = [[sb.StridedView(Bs[i], (m, m), (1, m)), sb.StridedView(As[i], (m, m), (m, 1))] for i in range(k)]
    ms = _time(lambda i: sb.run_mapreduce(prog, 0, 0, 0.0, (m, m), views[i % k]), 10 * k)
    out.append(_entry("C1 f64 1000^2 B .= 3 .* A' (rot)", 2 * m * m * 8, ms, peak, sets=k, kernel=_kernel(prog, 0, 0, (m, m), views[0])))
    ms = _time(lambda i: sb.run_mapreduce(prog, 0, 0, 0.0, (m, m), views[0]), 200)
    out.append(_entry("C1 f64 1000^2 B .= 3 .* A' (warm, L2-resident)", 2 * m * m * 8, ms, peak))
    del As, Bs, views

    # C3: F64 32^4 permutedims!(B, A, (4,3,2,1))
    for m in (32,):
        shape = (m,) * 4
        k = sets_for(2 * m ** 4 * 8)
        As = [torch.randn(m ** 4, dtype=torch.float64, device=dev) for _ in range(k)]
        Bs = [torch.empty(m ** 4, dtype=torch.float64, device=dev) for _ in range(k)]
        pairs = [(sb.StridedView(Bs[i], shape, _col(shape)), sb.StridedView(As[i], shape, _col(shape)).permutedims((3, 2, 1, 0))) for i in range(k)]
        ms = _time(lambda i: sb.copy_(*pairs[i % k]), 20 * k)
        out.append(_entry(f"C3 f64 {m}^4 permutedims (4,3,2,1) (rot)", 2 * m ** 4 * 8, ms, peak, sets=k, kernel=_kernel([], 0, 0, shape, list(pairs[0]))))
        ms = _time(lambda i: sb.copy_(*pairs[0]), 300)
        out.append(_entry(f"C3 f64 {m}^4 permutedims (4,3,2,1) (warm, L2-resident)", 2 * m ** 4 * 8, ms, peak))
        # independent problems of this size as ONE batch (sb_mapreduce_batch): calls that share a plan are merged into one
        # grouped launch (tma_kernel.cuh "GROUP"), so the tiles of all problems stream through one persistent grid instead of
        # paying one launch + one DRAM round trip per statement; time PER PROBLEM
        for nb in (8, 16):
            batches = [[([], 0, 0, 0.0, shape, list(pairs[(j * nb + q) % k])) for q in range(nb)] for j in range(max(1, k // nb))]
            eng.reset_stats()
            sb.run_batch(batches[0])
            st = eng.stats()
            ms = _time(lambda i: sb.run_batch(batches[i % len(batches)]), 10 * len(batches)) / nb
            out.append(_entry(f"C3 f64 {m}^4 permutedims (4,3,2,1) x{nb} in one batch (rot, per problem)", 2 * m ** 4 * 8, ms, peak, batch=nb,
                              launches_per_batch=st["launches"], grouped_calls=st["grouped_calls"],
                              kernel=f"sb_mapreduce_batch: {st['launches']} grouped launch(es) for {nb} calls (map_tile_group_kernel, the group's plan is sized for all "
                                     f"problems: 2048-element tiles, per-tile records); a single call of this shape runs " + _kernel([], 0, 0, shape, list(pairs[0]))))
        # the size-matched yardstick: a dense copy of the same 16.8 MB through the engine (what launch + ramp + drain cost
        # at this size, whatever the access pattern)
        dense = [(sb.StridedView(Bs[i]), sb.StridedView(As[i])) for i in range(k)]
        ms = _time(lambda i: sb.copy_(*dense[i % k]), 20 * k)
        out.append(_entry(f"dense copy f64 {m}^4 elements (rot; size-matched yardstick for C3 / C4')", 2 * m ** 4 * 8, ms, peak, sets=k,
                          kernel=_kernel([], 0, 0, dense[0][0].size, list(dense[0]))))
        del dense
        del As, Bs, pairs, batches

    # C4: F32 64^4 and C4': F64 32^4 4-way permutedims sum
    p4 = [A_(0), A_(1), CALL("add"), A_(2), CALL("add"), A_(3), CALL("add")]
    for m, dt, esz, nm in ((64, torch.float32, 4, "C4 f32 64^4"), (32, torch.float64, 8, "C4' f64 32^4")):
        shape = (m,) * 4
        k = sets_for(2 * m ** 4 * esz)
        As = [torch.randn(m ** 4, dtype=dt, device=dev) for _ in range(k)]
        Bs = [torch.empty(m ** 4, dtype=dt, device=dev) for _ in range(k)]
        vs = []
        for i in range(k):
            Av = sb.StridedView(As[i], shape, _col(shape))
            vs.append([sb.StridedView(Bs[i], shape, _col(shape))] + [Av.permutedims(p) for p in ((0, 1, 2, 3), (1, 2, 3, 0), (2, 3, 0, 1), (3, 0, 1, 2))])
        ms = _time(lambda i: sb.run_mapreduce(p4, 0, 0, 0.0, shape, vs[i % k]), 10 * k)
        out.append(_entry(f"{nm} 4-way permutedims sum (rot)", 2 * m ** 4 * esz, ms, peak, sets=k, operand_bytes=5 * m ** 4 * esz,
                          kernel=_kernel(p4, 0, 0, shape, vs[0])))
        ms = _time(lambda i: sb.run_mapreduce(p4, 0, 0, 0.0, shape, vs[0]), 100)
        out.append(_entry(f"{nm} 4-way permutedims sum (warm)", 2 * m ** 4 * esz, ms, peak))
        del As, Bs, vs

    # C5: F64 8x4096x4096 mapreduce(abs2, +, A; dims=(2,3)) on ONE GPU (1 GiB, larger than L2)
    g, kk = 8, 4096
    t = torch.randn(g * kk * kk, dtype=torch.float64, device=dev)
    o = torch.zeros(g, dtype=torch.float64, device=dev)
    T = sb.StridedView(t, (g, kk, kk), (1, g, g * kk))
    O = sb.StridedView(o, (g, kk, kk), (1, 0, 0))
    pa = [A_(0), CALL("abs2")]
    ms = _time(lambda i: sb.run_mapreduce(pa, 1, 0, 0.0, (g, kk, kk), [O, T]), 20)
    out.append(_entry("C5 f64 8x4096x4096 mapreduce(abs2,+;dims=(2,3)) 1 GPU", g * kk * kk * 8 + 64, ms, peak, kernel=_kernel(pa, 1, 0, (g, kk, kk), [O, T])))
    # one dense shard as each of 8 GPUs would hold it: 4096x4096 -> 1 scalar
    sh = torch.randn(kk * kk, dtype=torch.float64, device=dev)
    o1 = torch.zeros(1, dtype=torch.float64, device=dev)
    S = sb.StridedView(sh, (kk, kk), (1, kk))
    O1 = sb.StridedView(o1, (kk, kk), (0, 0))
    ms = _time(lambda i: sb.run_mapreduce(pa, 1, 0, 0.0, (kk, kk), [O1, S]), 50)
    out.append(_entry("C5 shard f64 4096x4096 -> scalar (per-GPU share of the 8-GPU run)", kk * kk * 8 + 8, ms, peak, kernel=_kernel(pa, 1, 0, (kk, kk), [O1, S])))
    eng.set_sync(True)
    return out
