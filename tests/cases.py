"""The parity matrix: the recipes of the reference's own test-suite (reference test/othertests.jl, restated
with NumPy-seeded data) plus the gaps SURVEY.md section 4 lists (negative strides, odd sizes, offsets,
Float32 reductions).  `scale` shrinks the sizes for the CPU emulation / oracle cross-checks; scale=1 is what
the reference uses and what the GPU tests run."""
from __future__ import annotations

import numpy as np

from helpers import (A, F, K, Case, ViewSpec, SEED, P_COPY, P_LAMBDA3, rand, randn, col_major_strides)

DTYPES = (np.float32, np.float64, np.complex64, np.complex128)


def _rng(*salt):
    # crc32, not hash(): str hashes are randomised per interpreter run, the fixtures of tests/golden/ must be reproducible
    import zlib
    return np.random.default_rng([SEED, *[zlib.crc32(repr(s).encode()) % (2 ** 31) for s in salt]])


# othertests.jl:1-15 "in-place matrix operations": conj!, adjoint!, transpose!, permutedims!(.., (2,1)); `==`
def inplace_matrix_cases(n=1000):
    out = []
    for dt in DTYPES:
        rng = _rng("inplace", np.dtype(dt).name)
        a1, a2 = randn(rng, n * n, dt), randn(rng, n * n, dt)
        V1, V2 = ViewSpec.dense(0, (n, n)), ViewSpec.dense(1, (n, n))
        name = np.dtype(dt).name
        if np.dtype(dt).kind == "c":
            out.append(Case(f"conj!_{name}", [a1], [V1, V1], [A(0), F("conj")]))
        out.append(Case(f"adjoint!_{name}", [a2, a1], [V1, ViewSpec(1, 0, (n, n), (n, 1), conj=True)], P_COPY))
        out.append(Case(f"transpose!_{name}", [a2, a1], [V1, ViewSpec(1, 0, (n, n), (n, 1))], P_COPY))
        out.append(Case(f"permutedims21_{name}", [a2, a1], [V1, ViewSpec.dense(1, (n, n)).permutedims((1, 0))], P_COPY))
    return out


# othertests.jl:17-44 "map, scale!, axpy! and axpby!": N = 2..6, dims div(60,N)^N, independent random perms
def map_axpy_cases(total=60, dtypes=DTYPES, Ns=(2, 3, 4, 5, 6)):
    out = []
    for dt in dtypes:
        for N in Ns:
            rng = _rng("map", np.dtype(dt).name, N)
            d = max(total // N, 2)
            shape = (d,) * N
            n = d ** N
            R = [rand(rng, n, dt) for _ in range(3)]
            B = [ViewSpec.dense(i, shape).permutedims(rng.permutation(N)) for i in range(3)]
            nm = f"{np.dtype(dt).name}_N{N}"
            out.append(Case(f"rmul_{nm}", R, [B[0], B[0]], [A(0), K(0.5), F("mul")]))
            out.append(Case(f"lmul_{nm}", R, [B[1], B[1]], [K(1 / 3), A(0), F("mul")]))
            out.append(Case(f"axpy_{nm}", R, [B[1], B[0], B[1]], [K(1 / 3), A(0), F("mul"), A(1), F("add")]))
            out.append(Case(f"axpby_{nm}", R, [B[2], B[0], B[2]],
                            [K(1 / 3), A(0), F("mul"), K(0.5), A(1), F("mul"), F("add")]))
            z = np.zeros(n, dt)
            out.append(Case(f"lambda3_{nm}", [z] + R, [ViewSpec.dense(0, shape)] + [ViewSpec(b.parent + 1, 0, b.size, b.strides) for b in B],
                            P_LAMBDA3, rtol=2e-5 if np.dtype(dt).itemsize <= 8 and np.dtype(dt).kind == "f" and dt == np.float32 else None))
            out.append(Case(f"mul_alpha_{nm}", R, [B[0], B[1]], [K(1), A(0), F("mul")]))
    return out


# othertests.jl:46-66 "broadcast with StridedView": mixed rank (10)x(10,10)x(10,10,10), adjoint, Ref scalar, max/abs/real
def broadcast_cases(n=10):
    out = []
    for dt in DTYPES:
        rng = _rng("bc", np.dtype(dt).name)
        R1, R2, R3 = rand(rng, n, dt), rand(rng, n * n, dt), rand(rng, n ** 3, dt)
        cplx = np.dtype(dt).kind == "c"
        B1 = ViewSpec(1, 0, (n, n, n), (1, 0, 0))
        B1_2d = ViewSpec(1, 0, (n, n), (1, 0))
        p2 = rng.permutation(2)
        B2 = ViewSpec.dense(2, (n, n)).permutedims(p2)
        B2adj = ViewSpec(2, 0, (n, n, n), (B2.strides[1], B2.strides[0], 0), conj=cplx)
        B3 = ViewSpec.dense(3, (n,) * 3).permutedims(rng.permutation(3))
        nm = np.dtype(dt).name
        z2, z3 = np.zeros(n * n, dt), np.zeros(n ** 3, dt)
        # B1 .+ sin.(B2 .- 3)
        out.append(Case(f"bc1_{nm}", [z2, R1, R2], [ViewSpec.dense(0, (n, n)), B1_2d, B2],
                        [A(0), A(1), K(3), F("sub"), F("sin"), F("add")], rtol=2e-5 if dt in (np.float32, np.complex64) else None))
        # B2' .* B3 .- Ref(0.5)
        out.append(Case(f"bc2_{nm}", [z3, R1, R2, R3], [ViewSpec.dense(0, (n,) * 3), B2adj, B3],
                        [A(0), A(1), F("mul"), K(0.5, typ=2), F("sub")]))
        # B2' .* B3 .- max.(abs.(B1), real.(B3))
        out.append(Case(f"bc3_{nm}", [z3, R1, R2, R3], [ViewSpec.dense(0, (n,) * 3), B2adj, B3, B1, B3],
                        [A(0), A(1), F("mul"), A(2), F("abs"), A(3), F("real"), F("max"), F("sub")]))
    return out


# othertests.jl:68-107 "mapreduce with StridedView": dims=(1,3,5) of 10^6, every initop flavour, (100,100,2)
def mapreduce_cases(n=10, n3=100):
    out = []
    for dt in DTYPES:
        rng = _rng("mr", np.dtype(dt).name)
        nm = np.dtype(dt).name
        full = (n,) * 6
        R1 = rand(rng, n ** 6, dt)
        V1 = ViewSpec.dense(1, full)
        tol = 2e-4 if dt in (np.float32, np.complex64) else 1e-11
        # sum(R1; dims=(1,3,5)) and mapreduce(sin, +, ...)
        ost = (0, 1, 0, n, 0, n * n)
        for fname, prog in (("sum", P_COPY), ("sin", [A(0), F("sin")])):
            out.append(Case(f"{fname}_135_{nm}", [np.zeros(n ** 3, dt), R1], [ViewSpec(0, 0, full, ost), V1], prog, op=1, rtol=tol))
        # direct _mapreducedim! with initop flavours, output sreshape(R2, (10,1,1,10,10,1))
        ost2 = (1, 0, 0, n, n * n, 0)
        beta = complex(rand(rng, 1, dt)[0])
        for iname, icode, ib in (("identity", 2, 0.0), ("zero", 1, 0.0), ("scale", 3, beta), ("const", 4, beta), ("conj", 5, 0.0), ("none", 0, 0.0)):
            R2 = rand(rng, n ** 3, dt)
            out.append(Case(f"initop_{iname}_{nm}", [R2, R1], [ViewSpec(0, 0, full, ost2), V1], [A(0), F("sin")], op=1,
                            initop=icode, init=ib, rtol=tol))
        # (100,100,2) regression: sum over dims (1,2)
        R3 = rand(rng, n3 * n3 * 2, dt)
        out.append(Case(f"sum_100_100_2_{nm}", [np.zeros(2, dt), R3],
                        [ViewSpec(0, 0, (n3, n3, 2), (0, 0, 1)), ViewSpec.dense(1, (n3, n3, 2))], P_COPY, op=1, rtol=tol))
    return out


# othertests.jl:109-128 "complete reductions": sum, maximum(abs), minimum(real), predicate count (==), permuted, prod(exp)
def complete_reduction_cases(n=10):
    out = []
    for dt in DTYPES:
        rng = _rng("cr", np.dtype(dt).name)
        nm = np.dtype(dt).name
        full = (n,) * 6
        R1 = rand(rng, n ** 6, dt) - (0.5 if np.dtype(dt).kind == "f" else 0.5 + 0.5j)
        R1 = R1.astype(dt)
        tol = 2e-4 if dt in (np.float32, np.complex64) else 1e-11
        rdt = np.float32 if dt in (np.float32, np.complex64) else np.float64
        for vname, V in (("dense", ViewSpec.dense(1, full)), ("perm", ViewSpec.dense(1, full).permutedims(rng.permutation(6)))):
            z = (0,) * 6
            out.append(Case(f"sum_all_{vname}_{nm}", [np.zeros(1, dt), R1], [ViewSpec(0, 0, full, z), V], P_COPY, op=1, rtol=tol))
            out.append(Case(f"max_abs_{vname}_{nm}", [np.zeros(1, rdt), R1], [ViewSpec(0, 0, full, z), V], [A(0), F("abs")], op=4))
            out.append(Case(f"min_real_{vname}_{nm}", [np.full(1, 9.0, rdt), R1], [ViewSpec(0, 0, full, z), V], [A(0), F("real")], op=3))
            out.append(Case(f"count_neg_{vname}_{nm}", [np.zeros(1, np.float64), R1], [ViewSpec(0, 0, full, z), V],
                            [A(0), F("real"), K(0), F("lt")], op=1, rtol=0.0))
        R3 = rand(rng, 125, dt)
        out.append(Case(f"prod_exp_{nm}", [np.ones(1, dt), R3], [ViewSpec(0, 0, (5, 5, 5), (0, 0, 0)), ViewSpec.dense(1, (5, 5, 5))],
                        [A(0), F("exp")], op=2, rtol=tol))
    return out


# Complete reductions over DENSE operands large enough for the streamed kernel (csrc/stream_kernel.cuh): ragged tails
# (element counts that are not a multiple of 16 bytes), 1-3 inputs (`sum(A)`, dot-like `sum(A .* B)`, benchmarks/
# benchtests.jl:44-68 `benchmark_sum`), every reduction operator, every eltype, a misaligned base (tile kernel at bind
# time), existing output contents participating (initop === nothing, src/mapreduce.jl:314).
def stream_reduction_cases(n=70001):
    out = []
    for dt in DTYPES:
        rng = _rng("stream", np.dtype(dt).name)
        nm = np.dtype(dt).name
        tol = 2e-4 if dt in (np.float32, np.complex64) else 1e-11
        x, y, z = (rand(rng, n + 5, dt) - 0.5).astype(dt), (rand(rng, n + 5, dt) - 0.5).astype(dt), (rand(rng, n + 5, dt) + 0.5).astype(dt)
        O = ViewSpec(0, 0, (n,), (0,))
        X, Y, Z = ViewSpec(1, 0, (n,), (1,)), ViewSpec(2, 0, (n,), (1,)), ViewSpec(3, 0, (n,), (1,))
        out.append(Case(f"stream_sum_{nm}", [np.full(1, 3, dt), x], [O, X], P_COPY, op=1, rtol=tol))
        out.append(Case(f"stream_sum_init0_{nm}", [np.full(1, 3, dt), x], [O, X], P_COPY, op=1, initop=1, rtol=tol))
        out.append(Case(f"stream_dot_{nm}", [np.zeros(1, dt), x, y], [O, X, Y], [A(0), A(1), F("mul")], op=1, rtol=tol))
        out.append(Case(f"stream_3in_{nm}", [np.zeros(1, dt), x, y, z], [O, X, Y, Z], [A(0), A(1), F("mul"), A(2), F("div")], op=1, rtol=tol * 50))
        out.append(Case(f"stream_sum_off1_{nm}", [np.zeros(1, dt), x], [O, ViewSpec(1, 1, (n,), (1,))], P_COPY, op=1, rtol=tol))
        out.append(Case(f"stream_sum_even_{nm}", [np.zeros(1, dt), x], [ViewSpec(0, 0, (n - 1,), (0,)), ViewSpec(1, 0, (n - 1,), (1,))], P_COPY, op=1, rtol=tol))
        # a FEW outputs, each reducing one dense run: mapreduce(f, op, A; dims=(1,2)) of a 3-D array (config 5 with several
        # dense slices per GPU), also through a permuted output, a stepped kept dim, and two kept dims
        m1, m2, g = (n // 3) // 2 * 2, 3, 5
        a3 = (rand(rng, m1 * m2 * g * 2, dt) - 0.5).astype(dt)
        A3 = ViewSpec(1, 0, (m1, m2, g), (1, m1, m1 * m2))
        out.append(Case(f"stream_dims12_{nm}", [np.full(g, 2, dt), a3], [ViewSpec(0, 0, (m1, m2, g), (0, 0, 1)), A3], P_COPY, op=1, rtol=tol))
        out.append(Case(f"stream_dims12_init0_rev_{nm}", [np.full(g, 2, dt), a3], [ViewSpec(0, g - 1, (m1, m2, g), (0, 0, -1)), A3], P_COPY,
                        op=1, initop=1, rtol=tol))
        out.append(Case(f"stream_dims12_step2_{nm}", [np.zeros(g, dt), a3], [ViewSpec(0, 0, (m1, m2, g), (0, 0, 1)),
                                                                             ViewSpec(1, 0, (m1, m2, g), (1, m1, 2 * m1 * m2))], P_COPY, op=1, rtol=tol))
        A4 = ViewSpec(1, 0, (m1 * m2 // 2, 2, g), (1, m1 * m2 // 2, m1 * m2))  # kept dims 2 x g, runs of m1*m2/2
        out.append(Case(f"stream_2kept_{nm}", [np.zeros(2 * g, dt), a3], [ViewSpec(0, 0, (m1 * m2 // 2, 2, g), (0, g, 1)), A4], P_COPY, op=1, rtol=tol))
        # kept dim INNERMOST and contiguous (column-major `mapreduce(f, op, A; dims=(2,3))`, BASELINE config 5 on one GPU):
        # the interleaved mode of the streamed kernel, K * sizeof(T) = 16 ... 512 bytes, ragged reduced extent
        esz = np.dtype(dt).itemsize
        for K in sorted({16 // esz, 64 // esz, min(512 // esz, 64)}):
            if K < 1:
                continue
            mm = (n * 8) // (K * esz) + 3
            ai = (rand(rng, K * mm, dt) - 0.5).astype(dt)
            AI = ViewSpec(1, 0, (K, mm), (1, K))
            out.append(Case(f"stream_inter{K}_{nm}", [np.full(K, 2, dt), ai], [ViewSpec(0, 0, (K, mm), (1, 0)), AI], P_COPY, op=1, rtol=tol))
            out.append(Case(f"stream_inter{K}_init0_rev_{nm}", [np.full(K, 2, dt), ai], [ViewSpec(0, K - 1, (K, mm), (-1, 0)), AI], P_COPY,
                            op=1, initop=1, rtol=tol))
        K = 64 // esz
        mm = (n * 8) // (K * esz) + 1
        ai, bi = (rand(rng, K * mm, dt) - 0.5).astype(dt), (rand(rng, K * mm, dt) - 0.5).astype(dt)
        out.append(Case(f"stream_inter_dot_{nm}", [np.zeros(K, dt), ai, bi], [ViewSpec(0, 0, (K, mm), (1, 0)), ViewSpec(1, 0, (K, mm), (1, K)),
                                                                                ViewSpec(2, 0, (K, mm), (1, K))], [A(0), A(1), F("mul")], op=1, rtol=tol))
        # 3-D form as the adapters pass it: dims (K, m1, m2), reduced dims fuse (src/mapreduce.jl:98-117)
        out.append(Case(f"stream_inter_3d_{nm}", [np.zeros(K, dt), ai], [ViewSpec(0, 0, (K, mm // 4, 4), (1, 0, 0)),
                                                                           ViewSpec(1, 0, (K, mm // 4, 4), (1, K, K * (mm // 4)))], P_COPY, op=1, rtol=tol))
        if np.dtype(dt).kind == "f":
            out.append(Case(f"stream_inter_max_{nm}", [np.full(K, -9, dt), ai], [ViewSpec(0, 0, (K, mm), (1, 0)), ViewSpec(1, 0, (K, mm), (1, K))],
                            [A(0), F("abs")], op=4))
            out.append(Case(f"stream_abs2_{nm}", [np.zeros(1, dt), x], [O, X], [A(0), F("abs2")], op=1, rtol=tol))
            out.append(Case(f"stream_dims12_max_{nm}", [np.full(g, -9, dt), a3], [ViewSpec(0, 0, (m1, m2, g), (0, 0, 1)), A3], [A(0), F("abs")], op=4))
            out.append(Case(f"stream_max_{nm}", [np.full(1, -9, dt), x], [O, X], P_COPY, op=4))
            out.append(Case(f"stream_min_{nm}", [np.full(1, 9, dt), x], [O, X], P_COPY, op=3))
            small = (1 + (rand(rng, 4100 * (8 // np.dtype(dt).itemsize) * 2, dt) - 0.5) * 1e-3).astype(dt)
            m = small.size
            out.append(Case(f"stream_prod_{nm}", [np.ones(1, dt), small], [ViewSpec(0, 0, (m,), (0,)), ViewSpec(1, 0, (m,), (1,))], P_COPY, op=2, rtol=tol))
    return out


# othertests.jl:130-190 "@strided macro": stepped ranges, views of adjoints, size-1 broadcast dims, reshape
def view_cases():
    out = []
    for dt in (np.float32, np.float64, np.complex128):
        rng = _rng("views", np.dtype(dt).name)
        nm = np.dtype(dt).name
        n = 20
        Apar, Bpar = randn(rng, n * n, dt), randn(rng, n * n, dt)
        # B[1:2:10, 3:7]-like window  =  A'[..] .+ 1
        out.append(Case(f"stepped_{nm}", [Bpar, Apar],
                        [ViewSpec(0, 2 * n + 1, (5, 4), (2, n)), ViewSpec(1, 3 + 4 * n, (5, 4), (3 * n, 2))],
                        [A(0), K(1), F("add")]))
        # negative strides on both sides (reverse range), SURVEY.md section 4 gap
        out.append(Case(f"negstride_{nm}", [Bpar, Apar],
                        [ViewSpec(0, n * n - 1, (n, n), (-1, -n)), ViewSpec(1, n - 1, (n, n), (-1, n))], [K(2), A(0), F("mul")]))
        # size-1 broadcast dim `A[4:4, :]`-like row broadcast over the columns of dest
        out.append(Case(f"rowbcast_{nm}", [Bpar, Apar], [ViewSpec.dense(0, (n, n)), ViewSpec(1, 3, (n, n), (0, n)), ViewSpec(1, 0, (n, n), (1, 0))],
                        [A(0), A(1), F("mul")]))
        # odd sizes + non-16B-aligned offsets
        m = 103
        P, Q = randn(rng, m * m + 7, dt), randn(rng, m * m + 7, dt)
        out.append(Case(f"odd103_{nm}", [Q, P], [ViewSpec(0, 3, (m, m), (1, m)), ViewSpec(1, 5, (m, m), (m, 1))], P_COPY))
        # sreshape of a permuted view: (6,6,5,4)-style
        S = randn(rng, 40 * 40, dt)
        Tt = np.zeros(36 * 20, dt)
        out.append(Case(f"reshaped_view_{nm}", [Tt, S], [ViewSpec.dense(0, (6, 6, 5, 4)), ViewSpec(1, 0, (6, 6, 5, 4), (1, 6, 40, 200))], P_COPY))
    # mixed eltypes: Float32 source into a Float64 destination, real into complex
    rng = _rng("mixed")
    x32 = randn(rng, 64 * 48, np.float32)
    out.append(Case("convert_f32_f64", [np.zeros(64 * 48, np.float64), x32],
                    [ViewSpec.dense(0, (64, 48)), ViewSpec.dense(1, (48, 64)).permutedims((1, 0))], P_COPY))
    xr = randn(rng, 64 * 48, np.float64)
    out.append(Case("real_into_complex", [np.zeros(64 * 48, np.complex128), xr],
                    [ViewSpec.dense(0, (64, 48)), ViewSpec.dense(1, (64, 48))], [A(0), K(0, 1, typ=2), F("mul")]))
    # abs2 of a complex array reduced into a real output
    zc = randn(rng, 32 * 33, np.complex128)
    out.append(Case("abs2_complex_to_real", [np.zeros(33, np.float64), zc],
                    [ViewSpec(0, 0, (32, 33), (0, 1)), ViewSpec.dense(1, (32, 33))], [A(0), F("abs2")], op=1))
    return out


# reductions that stress the kernel's shapes: row / column / strided / Float32 at scale
def reduction_shape_cases(big=1):
    out = []
    rng = _rng("redshape")
    m, n = 300 * big + 7, 200 * big + 3
    for dt in (np.float32, np.float64):
        nm = np.dtype(dt).name
        X = randn(rng, m * n, dt)
        tol = 5e-4 if dt == np.float32 else 1e-11
        V = ViewSpec.dense(1, (m, n))
        out.append(Case(f"colsum_{nm}", [np.zeros(n, dt), X], [ViewSpec(0, 0, (m, n), (0, 1)), V], P_COPY, op=1, rtol=tol))
        out.append(Case(f"rowsum_{nm}", [np.zeros(m, dt), X], [ViewSpec(0, 0, (m, n), (1, 0)), V], P_COPY, op=1, rtol=tol))
        out.append(Case(f"rowsum_T_{nm}", [np.zeros(n, dt), X], [ViewSpec(0, 0, (n, m), (1, 0)), ViewSpec(1, 0, (n, m), (m, 1))], P_COPY, op=1, rtol=tol))
        out.append(Case(f"maxcol_{nm}", [np.full(n, -np.inf, dt), X], [ViewSpec(0, 0, (m, n), (0, 1)), V], P_COPY, op=4))
        out.append(Case(f"sumsq_all_{nm}", [np.zeros(1, dt), X], [ViewSpec(0, 0, (m, n), (0, 0)), V], [A(0), F("abs2")], op=1, rtol=tol))
        # dot-product style: two inputs
        Y = randn(rng, m * n, dt)
        out.append(Case(f"dot_{nm}", [np.zeros(1, dt), X, Y], [ViewSpec(0, 0, (m, n), (0, 0)), V, ViewSpec.dense(2, (m, n))],
                        [A(0), A(1), F("mul")], op=1, rtol=tol))
    # generic matmul as a 3-D reduction with initop (linalg.jl:130-162): C = beta*C + sum_k alpha*A[m,k]*B[k,n]
    mm, nn, kk = 37, 29, 41
    Am, Bm, Cm = randn(rng, mm * kk, np.float64), randn(rng, kk * nn, np.float64), randn(rng, mm * nn, np.float64)
    out.append(Case("matmul_initop", [Cm, Am, Bm],
                    [ViewSpec(0, 0, (mm, nn, kk), (1, mm, 0)), ViewSpec(1, 0, (mm, nn, kk), (1, 0, mm)), ViewSpec(2, 0, (mm, nn, kk), (0, kk, 1))],
                    [A(0), A(1), F("mul"), K(0.7), F("mul")], op=1, initop=3, init=-0.3, rtol=1e-11))
    return out


def edge_cases():
    """empty and degenerate inputs (mapreduce.jl:48, :88-91)."""
    out = []
    z = np.zeros(4, np.float64)
    x = np.arange(6, dtype=np.float64)
    out.append(Case("empty_map", [z.copy(), x], [ViewSpec(0, 0, (0, 3), (1, 1)), ViewSpec(1, 0, (0, 3), (1, 1))], P_COPY))
    out.append(Case("empty_reduce_initop", [np.full(3, 5.0), x], [ViewSpec(0, 0, (3, 0), (1, 0)), ViewSpec(1, 0, (3, 0), (1, 3))],
                    P_COPY, op=1, initop=3, init=2.0))
    out.append(Case("single_element", [z.copy(), x], [ViewSpec(0, 2, (1, 1), (1, 1)), ViewSpec(1, 3, (1, 1), (1, 1))], [K(3), A(0), F("mul")]))
    out.append(Case("rank0", [z.copy(), x], [ViewSpec(0, 1, (), ()), ViewSpec(1, 4, (), ())], P_COPY))
    out.append(Case("fill_const", [z.copy()], [ViewSpec(0, 0, (4,), (1,))], [K(7.5, typ=2)]))
    out.append(Case("inplace_op_no_reduce", [np.arange(6, dtype=np.float64), x], [ViewSpec(0, 0, (2, 3), (1, 2)), ViewSpec(1, 0, (2, 3), (3, 1))],
                    [A(0), F("abs2")], op=1, initop=3, init=0.5))
    return out


# Odd extents (the reference's own tests use div(60, N)^N and 103; benchmarks/benchtests.jl sweeps 2^(2:1.5:20)): tiles that
# do not divide the dims -- balanced / shifted / masked tiles --, padded parents whose strides ARE 16-byte multiples while
# the extents are not (a shifted last tile then starts off the 16-byte grid), in-place updates (no recompute allowed).
def odd_extent_cases():
    out = []
    for dt in (np.float64, np.float32, np.complex64):
        rng = _rng("odd", np.dtype(dt).name)
        nm = np.dtype(dt).name
        for (m, n, ld) in ((71, 64, 72), (37, 129, 40), (103, 50, 104)):
            a, b = randn(rng, ld * n + 8, dt), np.zeros(ld * n + 8, dt)
            out.append(Case(f"odd_padded_copy_{m}x{n}_{nm}", [b, a], [ViewSpec(0, 0, (m, n), (1, ld)), ViewSpec(1, 0, (m, n), (1, ld))], P_COPY))
            out.append(Case(f"odd_padded_off_{m}x{n}_{nm}", [b.copy(), a], [ViewSpec(0, 2, (m, n), (1, ld)), ViewSpec(1, 4, (m, n), (1, ld))],
                            [K(3), A(0), F("mul")]))
        for m in (41, 70, 91):
            a, b = randn(rng, m * m, dt), np.zeros(m * m, dt)
            out.append(Case(f"odd_transpose_{m}_{nm}", [b, a], [ViewSpec.dense(0, (m, m)), ViewSpec(1, 0, (m, m), (m, 1))], P_COPY))
            out.append(Case(f"odd_A_plus_At_{m}_{nm}", [b.copy(), a], [ViewSpec.dense(0, (m, m)), ViewSpec.dense(1, (m, m)), ViewSpec(1, 0, (m, m), (m, 1))],
                            [A(0), A(1), F("add"), K(2), F("div")]))
            # in place: B .= B .+ A' (every element exactly once)
            out.append(Case(f"odd_inplace_{m}_{nm}", [a.copy(), a], [ViewSpec.dense(0, (m, m)), ViewSpec.dense(0, (m, m)), ViewSpec(1, 0, (m, m), (m, 1))],
                            [A(0), A(1), F("add")]))
        for m in (13, 21):
            sh = (m,) * 4
            a, b = randn(rng, m ** 4, dt), np.zeros(m ** 4, dt)
            for perm in ((3, 2, 1, 0), (1, 2, 3, 0), (2, 3, 0, 1)):
                out.append(Case(f"odd_permute_{m}^4_{''.join(map(str, perm))}_{nm}", [b, a], [ViewSpec.dense(0, sh), ViewSpec.dense(1, sh).permutedims(perm)], P_COPY))
    return out


# Partial reductions with MANY outputs along the contiguous dim (`sum(A; dims=2)` of a column-major matrix, Base.mapreducedim!
# through src/mapreduce.jl:74-96): tiles with many outputs, split reduced dim, thread-per-output fold of the partials.
def many_output_reduction_cases(n=768):
    out = []
    for dt in (np.float64, np.float32, np.complex64):
        rng = _rng("manyout", np.dtype(dt).name)
        nm = np.dtype(dt).name
        tol = 2e-4 if dt in (np.float32, np.complex64) else 1e-11
        a = (rand(rng, n * n, dt) - 0.5).astype(dt)
        M = ViewSpec.dense(1, (n, n))
        out.append(Case(f"rowsum_{n}_{nm}", [np.full(n, 2, dt), a], [ViewSpec(0, 0, (n, n), (1, 0)), M], P_COPY, op=1, rtol=tol))
        out.append(Case(f"rowsum_init0_abs2_{n}_{nm}", [np.full(n, 2, dt), a], [ViewSpec(0, 0, (n, n), (1, 0)), M], [A(0), F("abs2")], op=1, initop=1, rtol=tol))
        out.append(Case(f"colsum_{n}_{nm}", [np.zeros(n, dt), a], [ViewSpec(0, 0, (n, n), (0, 1)), M], P_COPY, op=1, rtol=tol))
        m = max(n // 8, 8)
        a3 = (rand(rng, m * m * m, dt) - 0.5).astype(dt)
        T = ViewSpec.dense(1, (m, m, m))
        out.append(Case(f"sum_dim2_{m}^3_{nm}", [np.zeros(m * m, dt), a3], [ViewSpec(0, 0, (m, m, m), (1, 0, m)), T], P_COPY, op=1, rtol=tol))
        out.append(Case(f"sum_dims23_{m}^3_{nm}", [np.zeros(m, dt), a3], [ViewSpec(0, 0, (m, m, m), (1, 0, 0)), T], P_COPY, op=1, rtol=tol))
        if np.dtype(dt).kind == "f":
            out.append(Case(f"rowmax_{n}_{nm}", [np.full(n, -9, dt), a], [ViewSpec(0, 0, (n, n), (1, 0)), M], [A(0), F("abs")], op=4))
    return out


# README.md:85-89 / :133-137: the compute-bound benchmark expression  B .= A .* exp.(-2 .* A) .+ sin.(A .* A)
# (the same parent captured four times: four identical views, nothing to fuse, one pass over A)
def readme_compute_bound_cases(n=1000):
    out = []
    for dt in (np.float64, np.float32):
        rng = _rng("readme_exp_sin", np.dtype(dt).name)
        a, b = randn(rng, n * n, dt), np.zeros(n * n, dt)
        V = ViewSpec.dense(1, (n, n))
        prog = [A(0), K(-2), A(1), F("mul"), F("exp"), F("mul"), A(2), A(3), F("mul"), F("sin"), F("add")]
        out.append(Case(f"readme_exp_sin_{np.dtype(dt).name}", [b, a], [ViewSpec.dense(0, (n, n)), V, V, V, V], prog,
                        rtol=2e-5 if dt == np.float32 else None))
    return out


def all_cases(scale=1.0):
    """scale < 1 shrinks the big recipes (for the CPU emulator); 1.0 = the reference's sizes."""
    s = scale
    cases = []
    cases += inplace_matrix_cases(max(int(1000 * s), 37))
    cases += map_axpy_cases(total=max(int(60 * s), 12) if s >= 1 else 20, Ns=(2, 3, 4, 5, 6))
    cases += broadcast_cases(10 if s >= 1 else 6)
    cases += mapreduce_cases(10 if s >= 1 else 4, 100 if s >= 1 else 23)
    cases += complete_reduction_cases(10 if s >= 1 else 4)
    cases += stream_reduction_cases(70001 if s >= 1 else 20011)
    cases += view_cases()
    cases += reduction_shape_cases(4 if s >= 1 else 1)
    cases += edge_cases()
    cases += odd_extent_cases()
    cases += many_output_reduction_cases(768 if s >= 1 else 160)
    cases += readme_compute_bound_cases(max(int(1000 * s), 64))
    return cases
