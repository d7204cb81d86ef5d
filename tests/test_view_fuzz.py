"""CPU suite: the StridedView layout contract (StridedViews.jl, external to the reference tree) as
mirrored by strided.jl_b200/view.py -- lazy permutedims / transpose / ranges with any step / integer indices /
sreshape -- fuzzed against NumPy on Fortran-ordered arrays.  Views are metadata only: `to_numpy()` is the check."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from helpers import sb


@st.composite
def view_ops(draw):
    n = draw(st.integers(1, 4))
    shape = tuple(draw(st.integers(1, 6)) for _ in range(n))
    ops = []
    cur = list(shape)
    for _ in range(draw(st.integers(0, 4))):
        kind = draw(st.sampled_from(["perm", "slice", "index", "reshape"])) if cur else "none"
        if kind == "perm" and len(cur) > 1:
            p = draw(st.permutations(list(range(len(cur)))))
            ops.append(("perm", tuple(p)))
            cur = [cur[i] for i in p]
        elif kind == "slice" and cur:
            idx = []
            new = []
            for s in cur:
                a = draw(st.integers(0, max(0, s - 1)))
                b = draw(st.integers(a, s))
                step = draw(st.sampled_from([1, 1, 2, 3, -1, -2]))
                sl = slice(a, b, step) if step > 0 else slice(b - 1 if b > 0 else None, a - 1 if a > 0 else None, step)
                idx.append(sl)
                new.append(len(range(*sl.indices(s))))
            ops.append(("slice", tuple(idx)))
            cur = new
        elif kind == "index" and len(cur) > 1:
            d = draw(st.integers(0, len(cur) - 1))
            if cur[d] == 0:
                continue
            k = draw(st.integers(-cur[d], cur[d] - 1))
            ops.append(("index", d, k))
            cur = cur[:d] + cur[d + 1:]
        elif kind == "reshape" and cur and all(c > 0 for c in cur):
            # merge two neighbouring dims or split one: only legal when expressible with strides (else DimensionMismatch)
            if len(cur) > 1 and draw(st.booleans()):
                d = draw(st.integers(0, len(cur) - 2))
                new = cur[:d] + [cur[d] * cur[d + 1]] + cur[d + 2:]
            else:
                d = draw(st.integers(0, len(cur) - 1))
                f = [q for q in (2, 3) if cur[d] % q == 0]
                if not f:
                    continue
                q = draw(st.sampled_from(f))
                new = cur[:d] + [q, cur[d] // q] + cur[d + 1:]
            ops.append(("reshape", tuple(new)))
            cur = new
    return shape, ops


@settings(max_examples=400, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(view_ops())
def test_lazy_view_operations_match_numpy(spec):
    shape, ops = spec
    n = int(np.prod(shape))
    flat = np.arange(n, dtype=np.float64) * 1.5 + 1
    ref = flat.reshape(shape, order="F")
    strides, acc = [], 1
    for s in shape:
        strides.append(acc)
        acc *= s
    v = sb.StridedView(flat, shape, tuple(strides))
    for op in ops:
        if op[0] == "perm":
            v, ref = v.permutedims(op[1]), np.transpose(ref, op[1])
        elif op[0] == "slice":
            v, ref = v[op[1]], ref[op[1]]
        elif op[0] == "index":
            idx = tuple(op[2] if d == op[1] else slice(None) for d in range(ref.ndim))
            v, ref = v[idx], ref[idx]
        else:
            try:
                v2 = v.sreshape(op[1])
            except sb.DimensionMismatch:
                break  # not expressible with strides: the reference's sreshape throws too (README.md "sreshape"); later ops assumed it
            v, ref = v2, np.reshape(ref, op[1], order="F")
        assert v.size == ref.shape
    got = v.to_numpy()
    assert got.shape == ref.shape and np.array_equal(got, ref)
