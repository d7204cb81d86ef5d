"""Alias-fused ("orbit") map path: every input is a dim-permuted view of ONE parent (`(A .+ A') ./ 2`, the 4-way
permutedims sum of README.md:91-104, reference src/mapreduce.jl:11-14 + src/broadcast.jl:27-37).

CPU part: the planner must pick the path for these shapes and the thread-grid emulation (TMA box load/store emulated,
real consumer body) must reproduce the NumPy semantic oracle bit for bit.  GPU part: the same cases plus the BASELINE
sizes through the C ABI, bit-exact against the oracle."""
import numpy as np
import pytest

from helpers import A, F, K, Case, ViewSpec, SEED, randn, case_c2, case_c4, P_SUM4


@pytest.fixture(autouse=True)
def _two_view_orbits(monkeypatch):
    """the fused path is the default from three aliased views on; two-view plans (A + A') take it on request"""
    monkeypatch.setenv("SB_ORBIT_NIN2", "1")


def _dense_pair(shape, dt, seed=SEED):
    rng = np.random.default_rng(seed)
    n = int(np.prod(shape))
    return randn(rng, n, dt), np.zeros(n, dt)


def orbit_cases(big=False):
    """(case, must_use_orbit); `big` adds the sizes of the reference's tests (1000 x 1000) for the GPU run"""
    s = lambda n: n  # noqa: E731
    out = []
    for dt in (np.float32, np.float64):
        nm = np.dtype(dt).name
        # C2 family: A + A' (edge tiles: sizes that are not multiples of 32)
        for n in (64, 260, 250) + ((1000,) if big else ()):
            c = case_c2(n, dt)
            c.name = f"orbit_avg_{nm}_{n}"
            out.append((c, (n * np.dtype(dt).itemsize) % 16 == 0 and n >= 260))
        # input 1 is itself the transposed view: B = A' + A (parent order != output order)
        n = s(96)
        a, b = _dense_pair((n, n), dt)
        Av = ViewSpec.dense(1, (n, n))
        out.append((Case(f"orbit_add2_tfirst_{nm}", [b, a], [ViewSpec.dense(0, (n, n)), Av.permutedims((1, 0)), Av],
                         [A(0), A(1), F("add")]), True))
        # axpby with a transposed alias: B = 2*A + 0.5*A'
        out.append((Case(f"orbit_axpby_{nm}", [b.copy(), a], [ViewSpec.dense(0, (n, n)), Av, Av.permutedims((1, 0))],
                         [K(2), A(0), F("mul"), K(0.5), A(1), F("mul"), F("add")]), True))
        # batched: B[i,j,z] = A[i,j,z] + A[j,i,z]  (unmoved batch dim)
        n, z = 48, 8
        a, b = _dense_pair((n, n, z), dt)
        Av = ViewSpec.dense(1, (n, n, z))
        out.append((Case(f"orbit_batched_{nm}", [b, a], [ViewSpec.dense(0, (n, n, z)), Av, Av.permutedims((1, 0, 2))],
                         [A(0), A(1), F("add")]), True))
        # 3-cycle: A[i,j,k] + A[j,k,i] + A[k,i,j]
        n = 32
        a, b = _dense_pair((n,) * 3, dt)
        Av = ViewSpec.dense(1, (n,) * 3)
        out.append((Case(f"orbit_sum3_{nm}", [b, a], [ViewSpec.dense(0, (n,) * 3), Av, Av.permutedims((1, 2, 0)), Av.permutedims((2, 0, 1))],
                         [A(0), A(1), F("add"), A(2), F("add")]), True))
        # a non-recipe element function over three aliased views: x * y' - z''  (in-kernel interpreter inside the fused kernel)
        out.append((Case(f"orbit_interp3_{nm}", [b.copy(), a], [ViewSpec.dense(0, (n,) * 3), Av, Av.permutedims((1, 2, 0)), Av.permutedims((2, 0, 1))],
                         [A(0), A(1), F("mul"), A(2), F("sub")]), True))
        # C4 family (edge tiles for 20)
        for n in (16, 20):
            c = case_c4(n, dt)
            c.name = f"orbit_sum4_{nm}_{n}"
            out.append((c, True))
        # 4-D pair swap: A[i,j,k,l] + A[j,i,l,k]  (two independent 2-cycles)
        n = 16
        a, b = _dense_pair((n,) * 4, dt)
        Av = ViewSpec.dense(1, (n,) * 4)
        out.append((Case(f"orbit_pairswap_{nm}", [b, a], [ViewSpec.dense(0, (n,) * 4), Av, Av.permutedims((1, 0, 3, 2))],
                         [A(0), A(1), F("add")]), True))
    return out


_CPU = orbit_cases()


@pytest.mark.parametrize("case,must", _CPU, ids=[c.name for c, _ in _CPU])
def test_orbit_emulated(case, must):
    plan = case.plan()
    if must:
        assert "orbit" in plan, plan
        o = plan["orbit"]
        assert o["ept"] * o["threads"] in (1024, 2048, 4096) and o["gmax"] <= 4 and o["smem_bytes"] <= 224 * 1024
    want = case.expected()
    case.assert_close(case.run_emul(), want, exact=True)
    case.assert_close(case.run_emul(grid_limit=2), want, exact=True)  # few persistent CTAs: ring wrap-around, staging parity


@pytest.mark.parametrize("logt", ["8", "9"])
def test_orbit_emulated_both_thread_counts(monkeypatch, logt):
    monkeypatch.setenv("SB_ORBIT_LOGT", logt)
    for c in (case_c4(16), case_c4(16, np.float64), case_c2(256), case_c2(128, np.float32)):
        assert c.plan()["orbit"]["threads"] == 1 << int(logt), c.name
        c.assert_close(c.run_emul(), exact=True)
        c.assert_close(c.run_emul(grid_limit=1), exact=True)


def test_orbit_not_chosen_when_output_aliases_parent_or_rows_unaligned():
    # in-place A .= A + A' is undefined for a parallel engine; the plan must not fuse it
    n = 64
    rng = np.random.default_rng(SEED)
    a = randn(rng, n * n, np.float64)
    Av = ViewSpec.dense(0, (n, n))
    c = Case("inplace", [a], [Av, Av, Av.permutedims((1, 0))], [A(0), A(1), F("add")])
    assert "orbit" not in c.plan()
    # row pitch 97*8 bytes is not a multiple of 16: not TMA-describable
    assert "orbit" not in case_c2(97).plan()
    # the BASELINE shapes do use it
    for c in (case_c2(4000), case_c4(64), case_c4(32, np.float64)):
        assert "orbit" in c.plan(), c.name


def test_two_view_plans_default_to_the_tma_ring(monkeypatch):
    monkeypatch.delenv("SB_ORBIT_NIN2")
    p = case_c2(4000).plan()
    assert "orbit" not in p and p["tma"] >= 2
    assert "orbit" in case_c4(64).plan()


_GPU = orbit_cases(big=True)


@pytest.mark.gpu
@pytest.mark.parametrize("case,must", _GPU, ids=[c.name for c, _ in _GPU])
def test_orbit_gpu(case, must):
    from strided_jl_b200.engine import Engine
    eng = Engine(0)  # fresh plan cache: the plan is built under this module's environment
    try:
        if must:
            assert "orbit" in case.plan()
        case.assert_close(case.run_gpu("device", engine=eng), exact=True)
    finally:
        eng.close()


@pytest.mark.gpu
def test_orbit_gpu_tma_store_and_direct_store_agree(monkeypatch):
    """both write-back modes of the fused kernel give the oracle's bytes (C4 has no edge tiles: direct store is legal)"""
    from strided_jl_b200.engine import Engine
    for direct in (False, True):
        if direct:
            monkeypatch.setenv("SB_ORBIT_DIRECT", "1")
        for c in (case_c4(32), case_c2(1024)):
            eng = Engine(0)
            try:
                assert c.plan()["orbit"]["direct_store"] == (1 if direct else 0)
                c.assert_close(c.run_gpu("device", engine=eng), exact=True)
            finally:
                eng.close()


def _random_family_case(rng, dt):
    """random rank, random set of 2..4 dim permutations of one parent (the first view may itself be permuted)"""
    N = int(rng.integers(2, 5))
    n = {2: int(rng.choice([64, 96, 128])), 3: int(rng.choice([16, 32])), 4: int(rng.choice([8, 16]))}[N]
    shape = (n,) * N
    nviews = int(rng.integers(2, 5))
    perms = []
    while len(perms) < nviews:
        p = tuple(int(x) for x in rng.permutation(N))
        if p not in perms:
            perms.append(p)
        if len(perms) == 1 and N == 2 and nviews > 2:
            nviews = 2  # only two distinct permutations of two dims
    a, b = _dense_pair(shape, dt, seed=int(rng.integers(1 << 30)))
    Av = ViewSpec.dense(1, shape)
    prog = {2: [A(0), A(1), F("add")], 3: [A(0), A(1), F("add"), A(2), F("add")], 4: P_SUM4}[len(perms)]
    return Case(f"orbit_random_N{N}_{'_'.join(''.join(map(str, p)) for p in perms)}_{np.dtype(dt).name}", [b, a],
                [ViewSpec.dense(0, shape)] + [Av.permutedims(p) for p in perms], prog)


def test_orbit_random_permutation_families_emulated():
    """Whatever the planner decides for a random family of aliased views (fused orbits when the generated group has
    orbits of <= 4 tiles and the output's fastest dim is moved, the generic / TMA kernels otherwise), the emulated
    kernels must reproduce the oracle bit for bit; the fused path must be hit for a fair share of the draws."""
    rng = np.random.default_rng(2026)
    fused = 0
    for i in range(40):
        c = _random_family_case(rng, (np.float32, np.float64)[i % 2])
        plan = c.plan()
        fused += "orbit" in plan
        c.assert_close(c.run_emul(), exact=True)
    assert fused >= 8, fused
