"""CPU suite: the host-side mirror of the reference interface (StridedView metadata, broadcast capture)."""
import numpy as np
import pytest

from helpers import sb, A, F, K


def _v(shape, dt=np.float64):
    n = int(np.prod(shape))
    return sb.StridedView(np.arange(n, dtype=dt), shape, tuple(int(np.prod(shape[:i])) for i in range(len(shape))))


def test_lazy_views_are_metadata_only():
    a = _v((3, 4, 5))
    p = a.permutedims((2, 0, 1))
    assert p.size == (5, 3, 4) and p.strides == (12, 1, 3) and p.offset == 0
    np.testing.assert_array_equal(p.to_numpy(), np.transpose(a.to_numpy(), (2, 0, 1)))
    s = a[1:3, ::2, 4]
    assert s.size == (2, 2) and s.strides == (1, 6) and s.offset == 1 + 48
    r = a[::-1, :, :]
    assert r.strides[0] == -1 and r.offset == 2
    np.testing.assert_array_equal(r.to_numpy(), a.to_numpy()[::-1])
    m = _v((4, 6), np.complex128)
    assert m.adjoint().conj_flag and m.adjoint().size == (6, 4) and not m.adjoint().adjoint().conj_flag


def test_sreshape_contract():
    a = _v((6, 6, 5, 4))
    assert a.sreshape((36, 20)).strides == (1, 36)
    assert a.sreshape((6, 3, 2, 5, 4)).strides == (1, 6, 18, 36, 180)
    v = sb.StridedView(np.zeros(1600), (36, 20), (1, 40))
    assert v.sreshape((6, 6, 5, 4)).strides == (1, 6, 40, 200)  # README.md example that IS strided
    with pytest.raises(sb.DimensionMismatch):
        v.sreshape((6, 3, 10, 4))  # README.md example that is NOT strided
    assert _v((10,)).sreshape((1, 10, 1)).size == (1, 10, 1)


def test_capture_order_and_program():
    A_ = _v((4, 4))
    B_ = _v((4, 4))
    bc = (A_ + A_.T) / 2
    assert [v.strides for v in sb.capturestridedargs(bc)] == [(1, 4), (4, 1)]
    assert sb.make_program(bc) == [A(0), A(1), F("add"), K(2), F("div")]
    bc = 3 * A_.T
    assert sb.make_program(bc) == [K(3), A(0), F("mul")]
    bc = ((A_ + B_) + A_.T) + B_.T
    assert sb.make_program(bc) == [A(0), A(1), F("add"), A(2), F("add"), A(3), F("add")]
    prog = sb.trace(lambda x, y, z: sb.sin(x) + y / sb.exp(-sb.abs_(z)), 3)
    assert prog == [A(0), F("sin"), A(1), A(2), F("abs"), F("neg"), F("exp"), F("div"), F("add")]
    assert sb.make_program(A_ - sb.Ref(0.5))[1] == K(0.5, typ=2)


def test_promoteshape_and_dimension_mismatch():
    v = _v((10,))
    (p,) = sb.promoteshape((10, 10, 10), v)
    assert p.size == (10, 10, 10) and p.strides == (1, 0, 0)
    with pytest.raises(sb.DimensionMismatch):
        sb.promoteshape((10, 7), _v((3, 7)))
    with pytest.raises(sb.DimensionMismatch):
        sb.map_("identity", _v((3, 4)), _v((4, 3)))


def test_opaque_callables_are_unsupported():
    with pytest.raises(sb.UnsupportedError):
        sb.trace(lambda x: np.sin(np.asarray(x)), 1)
    with pytest.raises(sb.UnsupportedError):
        sb.Broadcasted("erf", (1,))
