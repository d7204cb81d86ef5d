"""GPU suite: sb_mapreduce_batch -- a block of `_mapreduce_fuse!` calls issued as one batch (independent map calls overlap on
side streams, everything else runs in order).  The results must be those of the same calls made one after the other."""
import numpy as np
import pytest

from helpers import A, F, K, P_COPY, sb

pytestmark = pytest.mark.gpu


def _col(shape):
    st, acc = [], 1
    for s in shape:
        st.append(acc)
        acc *= s
    return tuple(st)


def _calls(dev, seed=3):
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    calls, checks = [], []
    sh = (16, 16, 16, 16)
    for i, p in enumerate([(3, 2, 1, 0), (1, 2, 3, 0), (2, 3, 0, 1), (0, 2, 1, 3), (3, 0, 2, 1), (1, 0, 3, 2)]):  # six independent permutes
        a = torch.randn(16 ** 4, dtype=torch.float64, device=dev, generator=g)
        b = torch.zeros_like(a)
        calls.append((P_COPY, 0, 0, 0.0, sh, [sb.StridedView(b, sh, _col(sh)), sb.StridedView(a, sh, _col(sh)).permutedims(p)]))
        q = tuple(3 - p[3 - d] for d in range(4))
        checks.append((b, a.view(*sh).permute(*q).contiguous().view(-1)))
    n = 300
    x = torch.randn(n * n, dtype=torch.float64, device=dev, generator=g)
    y, z = torch.zeros_like(x), torch.zeros_like(x)
    X, Y, Z = (sb.StridedView(t, (n, n), (1, n)) for t in (x, y, z))
    calls.append(([K(3), A(0), F("mul")], 0, 0, 0.0, (n, n), [Y, X.T]))             # Y = 3 X'          } a chain: must run in order
    calls.append(([A(0), A(1), F("add")], 0, 0, 0.0, (n, n), [Z, Y, X]))            # Z = Y + X         }
    checks.append((y, (3 * x.view(n, n)).t().contiguous().view(-1)))  # column-major flat of 3 X' == row-major flat of ... checked below via numpy
    s = torch.zeros(1, dtype=torch.float64, device=dev)
    calls.append(([A(0), F("abs2")], 1, 1, 0.0, (n, n), [sb.StridedView(s, (n, n), (0, 0)), X]))  # a reduction: after the join
    return calls, checks, (x, y, z, s, n)


def test_batch_equals_sequential():
    import torch
    dev = torch.device("cuda", 0)
    eng = sb.get_engine(0)
    calls, checks, (x, y, z, s, n) = _calls(dev)
    eng.reset_stats()
    sb.run_batch(calls)
    torch.cuda.synchronize()
    assert eng.stats()["batches"] == 1 and eng.stats()["launches"] == len(calls)
    for got, want in checks[:6]:
        assert torch.equal(got, want)
    xm = x.cpu().numpy().reshape((n, n), order="F")
    assert np.array_equal(y.cpu().numpy().reshape((n, n), order="F"), 3 * xm.T)
    assert np.array_equal(z.cpu().numpy().reshape((n, n), order="F"), 3 * xm.T + xm)
    np.testing.assert_allclose(s.item(), float((xm ** 2).sum()), rtol=1e-12)
    # the same calls one by one give the same bits
    outs = [c[5][0].parent.clone() for c in calls]
    for c in calls:
        c[5][0].parent.zero_()
    for c in calls:
        sb.run_mapreduce(*c)
    torch.cuda.synchronize()
    for c, o in zip(calls, outs):
        assert torch.equal(c[5][0].parent, o)


def test_batch_in_cuda_graph():
    import torch
    dev = torch.device("cuda", 0)
    eng = sb.get_engine(0)
    calls, checks, _ = _calls(dev, seed=5)
    eng.set_sync(False)
    side = torch.cuda.Stream()
    try:
        with torch.cuda.stream(side):
            sb.run_batch(calls)  # plans, side streams and events exist before the capture
            torch.cuda.synchronize()
            for c in calls:
                c[5][0].parent.zero_()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                sb.run_batch(calls)
            g.replay()
            torch.cuda.synchronize()
        for got, want in checks[:6]:
            assert torch.equal(got, want)
    finally:
        eng.set_sync(True)


def _same_plan_calls(dev, nprob, seed=11):
    """nprob independent problems of ONE plan: `permutedims!(B_i, A_i, (4,3,2,1))` at 32^4 (BASELINE config 3) -- and the
    torch results they must equal bit for bit."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    sh = (32, 32, 32, 32)
    calls, checks = [], []
    for _ in range(nprob):
        a = torch.randn(32 ** 4, dtype=torch.float64, device=dev, generator=g)
        b = torch.zeros_like(a)
        calls.append((P_COPY, 0, 0, 0.0, sh, [sb.StridedView(b, sh, _col(sh)), sb.StridedView(a, sh, _col(sh)).permutedims((3, 2, 1, 0))]))
        checks.append((b, a.view(*sh).permute(3, 2, 1, 0).contiguous().view(-1)))
    return calls, checks


@pytest.mark.parametrize("nprob", [2, 8, 19])
def test_same_plan_calls_share_one_launch(nprob):
    # equal-shape statements are merged into grouped launches of up to 16 problems (tma_kernel.cuh "GROUP")
    import torch
    dev = torch.device("cuda", 0)
    eng = sb.get_engine(0)
    calls, checks = _same_plan_calls(dev, nprob)
    assert calls[0][5][0].parent.numel() == 32 ** 4
    eng.reset_stats()
    sb.run_batch(calls)
    torch.cuda.synchronize()
    st = eng.stats()
    assert st["grouped_calls"] == nprob and st["launches"] == (nprob + 15) // 16
    for got, want in checks:
        assert torch.equal(got, want)


def test_grouped_two_input_map_and_graph_replay():
    # (A_i .+ B_i') ./ 2 at 512^2, eight problems in one launch; replayed from a CUDA graph with fresh data
    import torch
    dev = torch.device("cuda", 0)
    eng = sb.get_engine(0)
    n, nprob = 512, 8
    g = torch.Generator(device=dev)
    g.manual_seed(23)
    prog = [A(0), A(1), F("add"), K(2), F("div")]
    As = [torch.randn(n * n, dtype=torch.float64, device=dev, generator=g) for _ in range(nprob)]
    Bs = [torch.randn(n * n, dtype=torch.float64, device=dev, generator=g) for _ in range(nprob)]
    Os = [torch.zeros(n * n, dtype=torch.float64, device=dev) for _ in range(nprob)]
    calls = [(prog, 0, 0, 0.0, (n, n), [sb.StridedView(o, (n, n), (1, n)), sb.StridedView(a, (n, n), (1, n)), sb.StridedView(b, (n, n), (1, n)).T])
             for o, a, b in zip(Os, As, Bs)]

    def want(i):  # column-major flat of (A + B') / 2  ==  row-major flat of its transpose
        return ((As[i].view(n, n) + Bs[i].view(n, n).t()) / 2).contiguous().view(-1)

    eng.reset_stats()
    sb.run_batch(calls)
    torch.cuda.synchronize()
    assert eng.stats()["grouped_calls"] == nprob and eng.stats()["launches"] == 1
    for i in range(nprob):
        assert torch.equal(Os[i], want(i))
    eng.set_sync(False)
    side = torch.cuda.Stream()
    try:
        with torch.cuda.stream(side):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                sb.run_batch(calls)
            for a in As:
                a.normal_(generator=g)
            for o in Os:
                o.zero_()
            gr.replay()
            torch.cuda.synchronize()
        for i in range(nprob):
            assert torch.equal(Os[i], want(i))
    finally:
        eng.set_sync(True)


def test_grouping_respects_dependencies_and_inplace():
    # B = permute(A); C = permute(B): same plan, but the second reads what the first writes -> not grouped, in order
    import torch
    dev = torch.device("cuda", 0)
    eng = sb.get_engine(0)
    sh = (32, 32, 32, 32)
    a = torch.randn(32 ** 4, dtype=torch.float64, device=dev)
    b, c = torch.zeros_like(a), torch.zeros_like(a)
    V = lambda t: sb.StridedView(t, sh, _col(sh))
    calls = [(P_COPY, 0, 0, 0.0, sh, [V(b), V(a).permutedims((3, 2, 1, 0))]), (P_COPY, 0, 0, 0.0, sh, [V(c), V(b).permutedims((3, 2, 1, 0))])]
    eng.reset_stats()
    sb.run_batch(calls)
    torch.cuda.synchronize()
    assert eng.stats()["grouped_calls"] == 0 and eng.stats()["launches"] == 2
    assert torch.equal(c, a)  # the reversal is an involution (benchmarks/benchtests.jl:40)


@pytest.mark.parametrize("dtname", ["float64", "float32", "complex64"])
def test_grouped_lsu_launch_odd_extents(dtname):
    # odd extents cannot take the TMA ring: same-plan calls are merged into one launch of the LSU kernel (map_tile_group_kernel)
    import torch
    dt = getattr(torch, dtname)
    dev = torch.device("cuda", 0)
    eng = sb.get_engine(0)
    m, nprob = 21, 5
    sh = (m,) * 4
    g = torch.Generator(device=dev)
    g.manual_seed(31)
    calls, checks = [], []
    for _ in range(nprob):
        a = torch.randn(m ** 4, dtype=dt, device=dev, generator=g) if not dt.is_complex else torch.view_as_complex(torch.randn(m ** 4, 2, dtype=torch.float32, device=dev, generator=g))
        b = torch.zeros_like(a)
        calls.append((P_COPY, 0, 0, 0.0, sh, [sb.StridedView(b, sh, _col(sh)), sb.StridedView(a, sh, _col(sh)).permutedims((3, 2, 1, 0))]))
        checks.append((b, a.view(*sh).permute(3, 2, 1, 0).contiguous().view(-1)))
    eng.reset_stats()
    sb.run_batch(calls)
    torch.cuda.synchronize()
    st = eng.stats()
    assert st["grouped_calls"] == nprob and st["launches"] == 1
    for got, want in checks:
        assert torch.equal(got, want)


def test_grouped_lsu_two_input_sum():
    # Z_i = X_i .+ Y_i' at 301^2 (odd: LSU kernel), six problems in one launch
    import torch
    dev = torch.device("cuda", 0)
    eng = sb.get_engine(0)
    n, nprob = 301, 6
    g = torch.Generator(device=dev)
    g.manual_seed(37)
    Xs = [torch.randn(n * n, dtype=torch.float64, device=dev, generator=g) for _ in range(nprob)]
    Ys = [torch.randn(n * n, dtype=torch.float64, device=dev, generator=g) for _ in range(nprob)]
    Zs = [torch.zeros(n * n, dtype=torch.float64, device=dev) for _ in range(nprob)]
    calls = [([A(0), A(1), F("add")], 0, 0, 0.0, (n, n), [sb.StridedView(z, (n, n), (1, n)), sb.StridedView(x, (n, n), (1, n)), sb.StridedView(y, (n, n), (1, n)).T])
             for x, y, z in zip(Xs, Ys, Zs)]
    eng.reset_stats()
    sb.run_batch(calls)
    torch.cuda.synchronize()
    assert eng.stats()["grouped_calls"] == nprob and eng.stats()["launches"] == 1
    for x, y, z in zip(Xs, Ys, Zs):
        assert torch.equal(z, (x.view(n, n) + y.view(n, n).t()).contiguous().view(-1))


def test_plan_table_budget_drops_the_cache(monkeypatch):
    # plan tables (tile orders, per-tile records) are a cache with a device-memory budget: over budget the whole plan cache is
    # dropped at the next API entry and rebuilt on demand -- results stay right, plans are simply built again
    import torch
    dev = torch.device("cuda", 0)
    eng = sb.get_engine(0)
    monkeypatch.setenv("SB_PLAN_TABLE_MB", "1")
    eng.reload_env()
    try:
        eng.reset_stats()
        for rep in range(2):
            for m in (64, 72, 80, 88):  # four shapes with 0.2 ... 0.7 MB of per-tile records and tile orders each: 1.7 MB together
                sh = (m,) * 4
                a = torch.randn(m ** 4, dtype=torch.float64, device=dev)
                b = torch.zeros_like(a)
                sb.copy_(sb.StridedView(b, sh, _col(sh)), sb.StridedView(a, sh, _col(sh)).permutedims((3, 2, 1, 0)))
                torch.cuda.synchronize()
                assert torch.equal(b, a.view(*sh).permute(3, 2, 1, 0).contiguous().view(-1))
        assert eng.stats()["plans_built"] > 4  # the second round had to plan again at least once
    finally:
        monkeypatch.delenv("SB_PLAN_TABLE_MB")
        eng.reload_env()
