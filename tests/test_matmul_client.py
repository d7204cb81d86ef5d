"""The generic matrix-multiplication client of the hot path: `__mul!(C, A, B, α, β)` = one `_mapreducedim!` call with
initop ∈ {zero, nothing, x->x*β} and f ∈ {*, (x,y)->x*y*α} (reference src/linalg.jl:130-162).  Restates
test/blasmultests.jl:1-98 and test/othertests.jl:253-333: every combination of identity / conj / transpose / adjoint
on C, A and B, with and without α, β; zero-size k.  CPU: the descriptor the host mirror builds runs through the C
restatement of the reference (oracle/); GPU: through the C ABI."""
import itertools

import numpy as np
import pytest

from helpers import sb, oref, SEED

OPS = ("identity", "conj", "transpose", "adjoint")


def _apply(view, op):
    return {"identity": view, "conj": view.conj(), "transpose": view.transpose(), "adjoint": view.adjoint()}[op]


def _np_apply(a, op):
    return {"identity": a, "conj": a.conj(), "transpose": a.T, "adjoint": a.conj().T}[op]


def _case(dt, op1, op2, op3, m=11, n=7, k=13, seed=SEED):
    """C (m x n as seen through op3), A (m x k through op1), B (k x n through op2): parents sized accordingly"""
    rng = np.random.default_rng([seed, OPS.index(op1), OPS.index(op2), OPS.index(op3)])

    def parent(rows, cols, op):
        r, c = (cols, rows) if op in ("transpose", "adjoint") else (rows, cols)
        x = rng.standard_normal((r, c))
        if np.dtype(dt).kind == "c":
            x = x + 1j * rng.standard_normal((r, c))
        return np.asfortranarray(x.astype(dt))
    return parent(m, k, op1), parent(k, n, op2), parent(m, n, op3)


def _views(Ap, Bp, Cp, op1, op2, op3, wrap):
    def sv(p):
        flat = wrap(p.reshape(-1, order="F").copy())
        return sb.StridedView(flat, p.shape, (1, p.shape[0])), flat
    (A, fa), (B, fb), (C, fc) = sv(Ap), sv(Bp), sv(Cp)
    return _apply(C, op3), _apply(A, op1), _apply(B, op2), fc


def _expected(Ap, Bp, Cp, op1, op2, op3, alpha, beta):
    Cv = _np_apply(Cp, op3)
    want_view = beta * Cv + alpha * (_np_apply(Ap, op1) @ _np_apply(Bp, op2))
    # write back through op3 (a conj view stores the conjugate)
    out = Cp.copy()
    if op3 == "identity":
        out[...] = want_view
    elif op3 == "conj":
        out[...] = want_view.conj()
    elif op3 == "transpose":
        out[...] = want_view.T
    else:
        out[...] = want_view.conj().T
    return out


COMBOS = list(itertools.product(OPS, OPS, OPS))


@pytest.mark.parametrize("dt", [np.float64, np.complex128, np.complex64])
def test_generic_mul_descriptor_through_the_restated_reference(dt):
    tol = 1e-4 if dt == np.complex64 else 1e-12
    for (op1, op2, op3), (alpha, beta) in zip(COMBOS, itertools.cycle([(1, 0), (1, 1), (0.7, 0), (1, -0.3), (0.7, 1), (-1.5, 0.25)])):
        Ap, Bp, Cp = _case(dt, op1, op2, op3)
        C, A, B, fc = _views(Ap, Bp, Cp, op1, op2, op3, wrap=lambda x: x)
        call = sb._mul_generic_call(C, A, B, alpha, beta)
        f, op, initop, dims, arrays = call
        from strided_jl_b200.mapreduce import _initop, _op_code
        ic, b = _initop(initop)
        views = sb.promoteshape(dims, *arrays)
        desc = sb.make_desc(sb.trace(f, 2), _op_code(op), ic, b, dims, views)
        oref.mapreduce(desc, 3)
        got = fc.reshape(Cp.shape, order="F")
        np.testing.assert_allclose(got, _expected(Ap, Bp, Cp, op1, op2, op3, alpha, beta), rtol=tol, atol=tol)


def test_generic_mul_shortcuts_and_errors():
    Ap, Bp, Cp = _case(np.float64, "identity", "identity", "identity")
    C, A, B, _ = _views(Ap, Bp, Cp, "identity", "identity", "identity", wrap=lambda x: x)
    assert sb._mul_generic_call(C, A, B, 0, 0.5) is None  # alpha == 0 -> rmul!(C, beta)   (linalg.jl:141-142)
    A0 = sb.StridedView(np.zeros(0), (11, 0), (1, 11))
    B0 = sb.StridedView(np.zeros(0), (0, 7), (1, 0))
    assert sb._mul_generic_call(C, A0, B0, 1, 0) is None   # k == 0 (othertests.jl:288-296)
    with pytest.raises(sb.DimensionMismatch):
        sb._mul_generic_call(C, B, A)


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float64, np.complex128, np.float32])
def test_generic_mul_gpu_all_op_combinations(dt):
    import torch
    tol = 2e-4 if dt == np.float32 else 1e-11
    for (op1, op2, op3), (alpha, beta) in zip(COMBOS, itertools.cycle([(1, 0), (1, 1), (0.7, 0), (1, -0.3), (0.7, 1), (-1.5, 0.25)])):
        Ap, Bp, Cp = _case(dt, op1, op2, op3, m=61, n=47, k=103)  # 103: the odd size blasmultests.jl uses
        C, A, B, fc = _views(Ap, Bp, Cp, op1, op2, op3, wrap=lambda x: torch.from_numpy(x).cuda())
        sb.mul_generic_(C, A, B, alpha, beta)
        torch.cuda.synchronize()
        got = fc.cpu().numpy().reshape(Cp.shape, order="F")
        np.testing.assert_allclose(got, _expected(Ap, Bp, Cp, op1, op2, op3, alpha, beta), rtol=tol, atol=tol * 10)


@pytest.mark.gpu
def test_generic_mul_gpu_zero_k_and_alpha_zero():
    import torch
    Ap, Bp, Cp = _case(np.float64, "identity", "identity", "identity")
    C, A, B, fc = _views(Ap, Bp, Cp, "identity", "identity", "identity", wrap=lambda x: torch.from_numpy(x).cuda())
    sb.mul_generic_(C, A, B, 0, 0.5)
    torch.cuda.synchronize()
    np.testing.assert_allclose(fc.cpu().numpy().reshape(Cp.shape, order="F"), 0.5 * Cp)
