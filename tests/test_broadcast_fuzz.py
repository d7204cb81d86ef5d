"""CPU suite: the broadcast front end of the host mirror -- `capturestridedargs`, `promoteshape`, `make_program`
(reference src/broadcast.jl:27-98) -- fuzzed end to end: random expression trees over StridedViews of different ranks
and size-1 dims, lowered exactly as `materialize_` does and executed by the restated reference CPU path (oracle/),
compared with NumPy broadcasting.  (The device path runs the same descriptors; tests/test_gpu_parity.py.)"""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from helpers import sb, oref


def _col(shape):
    out, acc = [], 1
    for s in shape:
        out.append(acc)
        acc *= s
    return tuple(out)


@st.composite
def exprs(draw):
    n = draw(st.integers(1, 4))
    full = tuple(draw(st.integers(1, 5)) for _ in range(n))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    leaves = []
    for _ in range(draw(st.integers(1, 3))):
        r = draw(st.integers(1, n))  # lower rank = trailing dims missing (Julia broadcasting pads with 1 on the right)
        shape = tuple(full[d] if draw(st.booleans()) or full[d] == 1 else 1 for d in range(r))
        arr = np.asfortranarray(rng.standard_normal(shape) + 2.5)
        perm = tuple(draw(st.permutations(list(range(r))))) if draw(st.booleans()) else tuple(range(r))
        leaves.append((arr, perm))

    def build(depth):
        kind = draw(st.sampled_from(["leaf", "leaf", "un", "bin", "bin", "const"] if depth < 3 else ["leaf"]))
        if kind == "leaf":
            return ("leaf", draw(st.integers(0, len(leaves) - 1)))
        if kind == "const":
            return ("bin", draw(st.sampled_from(["add", "mul", "sub"])), ("leaf", draw(st.integers(0, len(leaves) - 1))),
                    ("const", draw(st.sampled_from([2, 0.5, -3]))))
        if kind == "un":
            return ("un", draw(st.sampled_from(["abs", "neg", "abs2", "exp"])), build(depth + 1))
        return ("bin", draw(st.sampled_from(["add", "sub", "mul", "div", "max"])), build(depth + 1), build(depth + 1))
    return full, leaves, build(0)


def _lower(tree, views):
    if tree[0] == "leaf":
        return views[tree[1]]
    if tree[0] == "const":
        return tree[1]
    if tree[0] == "un":
        return sb.Broadcasted(tree[1], (_lower(tree[2], views),))
    return sb.Broadcasted({"max": "max"}.get(tree[1], tree[1]), (_lower(tree[2], views), _lower(tree[3], views)))


def _numpy(tree, arrs):
    if tree[0] == "leaf":
        return arrs[tree[1]]
    if tree[0] == "const":
        return tree[1]
    if tree[0] == "un":
        x = _numpy(tree[2], arrs)
        return {"abs": np.abs, "neg": np.negative, "abs2": lambda v: v * v, "exp": np.exp}[tree[1]](x)
    a, b = _numpy(tree[2], arrs), _numpy(tree[3], arrs)
    return {"add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide, "max": np.maximum}[tree[1]](a, b)


@settings(max_examples=250, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(exprs())
def test_broadcast_lowering_matches_numpy_broadcasting(spec):
    full, leaves, tree = spec
    n = len(full)
    views, arrs = [], []
    for arr, perm in leaves:
        flat = arr.reshape(-1, order="F").copy()
        v = sb.StridedView(flat, arr.shape, _col(arr.shape))
        # a lazily permuted view of a permuted parent: same logical array, different strides
        if perm != tuple(range(arr.ndim)):
            inv = tuple(int(i) for i in np.argsort(perm))
            parent = np.asfortranarray(np.transpose(arr, perm))
            pf = parent.reshape(-1, order="F").copy()
            v = sb.StridedView(pf, parent.shape, _col(parent.shape)).permutedims(inv)
        views.append(v)
        arrs.append(arr.reshape(arr.shape + (1,) * (n - arr.ndim)))
    bc = _lower(tree, views)
    want = np.broadcast_to(_numpy(tree, arrs), full) if not isinstance(bc, sb.StridedView) else np.broadcast_to(arrs[tree[1]], full)
    if isinstance(bc, sb.StridedView):
        bc = sb.Broadcasted("identity", (bc,))
    dest_flat = np.zeros(int(np.prod(full)), np.float64)
    dest = sb.StridedView(dest_flat, full, _col(full))
    pviews = sb.promoteshape(dest.size, *sb.capturestridedargs(bc))
    try:
        desc = sb.make_desc(sb.make_program(bc), 0, 0, 0.0, dest.size, [dest] + list(pviews))
    except sb.UnsupportedError:
        return  # more than SB_MAX_OPS captured arguments: the glue falls back to the CPU method
    oref.mapreduce(desc, 2)
    got = dest_flat.reshape(full, order="F")
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)
