"""GPU suite (pytest -m gpu): the CUDA engine, called through the C ABI, against the oracles.

Bars (north_star): bit-exact for copy/permute/adjoint/conj and for single-rounding maps; <= 1e-6 relative for
fp64 reductions (measured ~1e-15); Float32 within a few ulp-scale tolerances written in cases.py."""
import numpy as np
import pytest

import cases
from helpers import case_c1, case_c2, case_c3, case_c4, case_c5, sb

pytestmark = pytest.mark.gpu

_TRANSCENDENTAL = {7, 8, 9, 10, 11, 12}


def _exact(case):
    return (case.op == 0 and not any(t[0] == 2 and t[1] in _TRANSCENDENTAL for t in case.tokens)
            and case.parents[case.views[0].parent].dtype.kind != "c")


@pytest.mark.parametrize("case", cases.all_cases(1.0), ids=lambda c: c.name)
def test_device_pointers(case):
    case.assert_close(case.run_gpu("device"), exact=_exact(case))


@pytest.mark.parametrize("case", cases.all_cases(0.3)[::3], ids=lambda c: c.name)
def test_host_pointers(case):
    case.assert_close(case.run_gpu("host"), exact=_exact(case))


def test_copy_permute_bit_exact_all_dtypes():
    for c in cases.inplace_matrix_cases(1000):
        c.assert_close(c.run_gpu("device"), exact=True)


def test_against_restated_reference():
    """same seeded inputs through the C restatement of the reference (4 tasks) and the CUDA path"""
    for c in (case_c1(1000), case_c2(1000), case_c3(32), case_c4(32)):
        assert c.run_gpu("device").tobytes() == c.run_ref(4).tobytes(), c.name
    c = case_c5(8, 512)
    np.testing.assert_allclose(c.run_gpu("device"), c.run_ref(1), rtol=1e-6)


def test_baseline_configs_full_size():
    import torch
    # C2: (A + A')/2 is symmetric and idempotent: f(f(A)) == f(A) bitwise; trace is preserved
    n = 4000
    rng = np.random.default_rng(1234)
    a = torch.from_numpy(rng.standard_normal(n * n)).cuda()
    b, b2 = torch.empty_like(a), torch.empty_like(a)
    A, B, B2 = (sb.StridedView(t, (n, n), (1, n)) for t in (a, b, b2))
    B.assign((A + A.T) / 2)
    B2.assign((B + B.T) / 2)
    torch.cuda.synchronize()
    bm = b.view(n, n)
    assert torch.equal(bm, bm.t()) and torch.equal(b, b2)
    assert torch.equal(torch.diagonal(bm), torch.diagonal(a.view(n, n)))
    ref = (a.view(n, n) + a.view(n, n).t()) * 0.5
    assert torch.equal(bm, ref)
    # C3: the reversal permutation is an involution
    m = 32
    x = torch.from_numpy(rng.standard_normal(m ** 4)).cuda()
    y, z = torch.empty_like(x), torch.empty_like(x)
    X, Y, Z = (sb.StridedView(t, (m,) * 4, (1, m, m * m, m ** 3)) for t in (x, y, z))
    sb.permutedims_(Y, X, (3, 2, 1, 0))
    sb.permutedims_(Z, Y, (3, 2, 1, 0))
    torch.cuda.synchronize()
    assert torch.equal(x, z)
    assert torch.equal(y.view(m, m, m, m), x.view(m, m, m, m).permute(3, 2, 1, 0).contiguous())
    # C4: Float32 64^4 4-way sum; invariant under a cyclic rotation of the OUTPUT indices
    c4 = case_c4(64)
    got = c4.run_gpu("device")
    c4.assert_close(got, exact=True)
    g4 = got.reshape((64,) * 4, order="F")
    a4 = c4.parents[1].reshape((64,) * 4, order="F")
    want = ((a4 + np.transpose(a4, (1, 2, 3, 0))) + np.transpose(a4, (2, 3, 0, 1))) + np.transpose(a4, (3, 0, 1, 2))
    assert np.array_equal(g4, want)
    # C5: Float64 8x4096x4096 abs2-sum over dims (2,3) vs a float64 torch reduction and vs math.fsum on a slice
    g, k = 8, 4096
    t = torch.from_numpy(rng.standard_normal(g * k * k)).cuda()
    out = torch.zeros(g, dtype=torch.float64, device="cuda")
    T = sb.StridedView(t, (g, k, k), (1, g, g * k))
    res = sb.mapreduce("abs2", "+", T, dims=(1, 2))
    torch.cuda.synchronize()
    want = (t.view(k * k, g) ** 2).sum(dim=0)
    got = torch.from_numpy(res.to_numpy().reshape(-1)).cuda()
    assert torch.allclose(got, want, rtol=1e-12, atol=0)
    import math
    exact0 = math.fsum((t.view(k * k, g)[:, 0].cpu().numpy() ** 2).tolist())
    assert abs(got[0].item() - exact0) / exact0 < 1e-12


def test_public_api_roundtrip_and_reductions():
    import torch
    rng = np.random.default_rng(7)
    a = torch.from_numpy(rng.standard_normal((50, 60, 7))).cuda()
    A = sb.StridedView(a)  # row-major torch tensor: strides (420, 7, 1)
    s = sb.sum(A)
    assert abs(s - a.sum().item()) < 1e-9
    assert sb.maximum(A, f="abs") == a.abs().max().item()
    assert sb.minimum(A) == a.min().item()
    r = sb.sum(A, dims=(0, 2))
    np.testing.assert_allclose(r.to_numpy().reshape(-1), a.sum(dim=(0, 2)).cpu().numpy(), rtol=1e-12)
    m = sb.map(lambda x, y: sb.sin(x) * y, A, A.permutedims((0, 1, 2)))
    np.testing.assert_allclose(m.to_numpy(), (torch.sin(a) * a).cpu().numpy(), rtol=1e-12)
    y = a.clone()
    sb.axpy_(0.5, A, sb.StridedView(y))
    assert torch.equal(y, 0.5 * a + a)


def test_engine_counts_launches_and_caches_plans():
    import torch
    eng = sb.get_engine(0)
    c = case_c2(256)
    c.run_gpu("device")
    eng.reset_stats()
    c.run_gpu("device")
    st = eng.stats()
    assert st["launches"] == 1 and st["plans_cached"] == 1 and st["plans_built"] == 0


def test_back_to_back_dependent_launches_are_ordered():
    """Every kernel is launched with programmatic stream serialisation (PDL): the next kernel's prologue overlaps the
    previous kernel's tail, and griddepcontrol.wait orders the data accesses.  A chain of DEPENDENT calls without any
    host synchronisation (each reads what the previous one wrote; the kernel families alternate) must equal NumPy."""
    import torch
    rng = np.random.default_rng(7)
    n = 1024
    a0 = rng.standard_normal(n * n)
    bufs = [torch.from_numpy(a0.copy()).cuda(), torch.zeros(n * n, dtype=torch.float64, device="cuda")]
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    ref = a0.reshape(n, n, order="F").copy()
    ref_acc = 0.0
    eng = sb.get_engine(0)
    eng.set_sync(False)
    try:
        for it in range(24):
            src, dst = bufs[it % 2], bufs[(it + 1) % 2]
            S, D = sb.StridedView(src, (n, n), (1, n)), sb.StridedView(dst, (n, n), (1, n))
            if it % 3 == 0:      # TMA ring kernel: D = (S + S') / 2
                D.assign((S + S.T) / 2)
                ref = (ref + ref.T) / 2
            elif it % 3 == 1:    # generic / TMA transpose-scale: D = 3 * S'
                D.assign(3 * S.T)
                ref = 3 * ref.T
            else:                # dense map + a full reduction of the fresh output into a running scalar
                D.assign(S * 0.25)
                ref = ref * 0.25
                sb.run_mapreduce([(0, 0, 0.0, 0.0), (2, sb.abi.FN["abs2"], 0.0, 0.0)], 1, 0, 0.0, (n, n),
                                 [sb.StridedView(acc, (n, n), (0, 0)), D])
                ref_acc += float(np.sum(ref * ref))
        torch.cuda.synchronize()
    finally:
        eng.set_sync(True)
    got = bufs[24 % 2].cpu().numpy().reshape(n, n, order="F")
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=0)
    np.testing.assert_allclose(acc.item(), ref_acc, rtol=1e-10)


def test_golden_fixtures_through_the_cuda_path():
    """tests/golden/*.npz (inputs + expected outputs committed with the script that made them) through the C ABI"""
    import os
    from helpers import Case, ROOT
    gdir = os.path.join(ROOT, "tests", "golden")
    files = sorted(f for f in os.listdir(gdir) if f.endswith(".npz"))
    assert len(files) >= 17
    for f in files:
        z = np.load(os.path.join(gdir, f), allow_pickle=False)
        case = Case.from_npz(z)
        case.assert_close(case.run_gpu("device"), z["expected"], exact=bool(z["exact"]))
