"""Base-interface adapters: Python mirror of reference src/mapreduce.jl:1-96 (and the BLAS-1 style wrappers of
src/linalg.jl:2-42, which are one-line broadcasts).  Every function ends in the same funnel as the
reference -- `_mapreduce_fuse!(f, op, initop, dims, arrays)` -- which here is `engine.run_mapreduce`.
Names keep Julia's `!` as a trailing underscore.
"""
from __future__ import annotations

from . import abi
from .broadcast import trace, promoteshape, result_dtype, Broadcasted, materialize_
from .engine import run_mapreduce, similar_parent
from .view import StridedView, DimensionMismatchError, maybestrided

_OPS = {"+": abi.SB_OP_ADD, "add": abi.SB_OP_ADD, "add_sum": abi.SB_OP_ADD, "*": abi.SB_OP_MUL, "mul": abi.SB_OP_MUL,
        "mul_prod": abi.SB_OP_MUL, "min": abi.SB_OP_MIN, "max": abi.SB_OP_MAX}


def _op_code(op):
    if isinstance(op, int):
        return op
    if op in _OPS:
        return _OPS[op]
    # reference: error("unknown reduction; incompatible with multithreading") (mapreduce.jl:188-190)
    raise abi.UnsupportedError(abi.SB_E_UNSUPPORTED, f"unknown reduction {op!r}")


def _initop(initop):
    """initop flavours (linalg.jl:145-158, othertests.jl:76-102) -> (code, beta)."""
    if initop is None:
        return abi.SB_INIT_NONE, 0.0
    if isinstance(initop, str):
        table = {"zero": abi.SB_INIT_ZERO, "identity": abi.SB_INIT_IDENTITY, "conj": abi.SB_INIT_CONJ}
        if initop in table:
            return table[initop], 0.0
    if isinstance(initop, tuple) and len(initop) == 2 and initop[0] in ("scale", "const"):
        return (abi.SB_INIT_SCALE if initop[0] == "scale" else abi.SB_INIT_CONST), initop[1]
    raise abi.UnsupportedError(abi.SB_E_UNSUPPORTED, f"initop {initop!r} is not expressible on the device path")


# ---- methods based on map!  (mapreduce.jl:2-14, 32-53) ---------------------------------------------------
def map_(f, b, a1, *A):
    b, a1 = maybestrided(b), maybestrided(a1)
    A = [maybestrided(a) for a in A]
    dims = b.size
    for a in (a1, *A):  # mapreduce.jl:43-46
        if a.size != dims:
            raise DimensionMismatchError(f"map!: size {a.size} != {dims}")
    if any(d == 0 for d in dims):  # :48
        return b
    tokens = trace(f, 1 + len(A))
    run_mapreduce(tokens, abi.SB_OP_NONE, abi.SB_INIT_NONE, 0.0, dims, [b, a1, *A])
    return b


def map(f, a1, *A):  # noqa: A001  (mirrors Base.map)
    a1 = maybestrided(a1)
    A = [maybestrided(a) for a in A]
    tokens = trace(f, 1 + len(A))
    out = similar_parent(a1, result_dtype(tokens, [a1, *A]), a1.size)
    return map_(tokens, out, a1, *A)


def copy_(dst, src):
    return map_("identity", dst, src)


def conj_(a):
    a = maybestrided(a)
    if not a.is_complex:  # conj!(a::StridedView{<:Real}) = a   (:5)
        return a
    return map_("conj", a, a)


def adjoint_(dst, src):
    return copy_(dst, maybestrided(src).adjoint())


def transpose_(dst, src):
    return copy_(dst, maybestrided(src).transpose())


def permutedims_(dst, src, p):
    return copy_(dst, maybestrided(src).permutedims(p))


# ---- reductions  (mapreduce.jl:16-30, 55-96) ---------------------------------------------------------------
def _mapreducedim_(f, op, initop, dims, arrays):
    """_mapreducedim!(f, op, initop, dims, arrays)   (mapreduce.jl:86-96)."""
    arrays = [maybestrided(a) for a in arrays]
    opc = _op_code(op)
    ic, beta = _initop(initop)
    tokens = trace(f, len(arrays) - 1)
    if any(d == 0 for d in dims):
        # :88-91 -- the engine applies initop to a non-empty output itself (abi.cu: empty_initop_desc)
        views = [arrays[0]._with(dims, _padded(arrays[0], dims))] + [a._with(dims, _padded(a, dims)) for a in arrays[1:]]
        run_mapreduce(tokens, opc, ic, beta, dims, views)
        return arrays[0]
    views = promoteshape(dims, *arrays)
    run_mapreduce(tokens, opc, ic, beta, dims, views)
    return arrays[0]


def _padded(a, dims):
    st = list(a.strides) + [1] * (len(dims) - a.ndim)
    sz = list(a.size) + [1] * (len(dims) - a.ndim)
    return [0 if (s == 1 and d != 1) else t for s, d, t in zip(sz, dims, st)]


def mapreducedim_(f, op, b, a1, *A):
    """Base.mapreducedim!(f, op, b, a1, A...)   (mapreduce.jl:74-84)."""
    b, a1 = maybestrided(b), maybestrided(a1)
    A = [maybestrided(a) for a in A]
    n = b.ndim
    if any(a.ndim != n for a in (a1, *A)):
        raise DimensionMismatchError("mapreducedim!: operands must have equal rank")
    dims = tuple(max(x) for x in zip(b.size, a1.size, *[a.size for a in A]))
    for a in (b, a1, *A):  # check_broadcast_axes (:81)
        for s, d in zip(a.size, dims):
            if s != d and s != 1:
                raise DimensionMismatchError(f"mapreducedim!: cannot broadcast size {a.size} to {dims}")
    return _mapreducedim_(f, op, None, dims, (b, a1, *A))


def _fill_scalar(out, value):
    materialize_(out, Broadcasted("identity", (value,)))


def _mapreduce_all(f, op, A, init=None):
    """_mapreduce (complete reduction)   (mapreduce.jl:55-72)."""
    A = maybestrided(A)
    tokens = trace(f, 1)
    opc = _op_code(op)
    T = result_dtype(tokens, [A])
    if len(A) == 0:  # Base.mapreduce_empty
        if opc == abi.SB_OP_ADD:
            b = 0.0
        elif opc == abi.SB_OP_MUL:
            b = 1.0
        else:
            raise ValueError("reducing over an empty collection is not allowed")
        return b if init is None else (b + init if opc == abi.SB_OP_ADD else b * init)
    out = similar_parent(A, T, (1,))
    ones = (1,) * A.ndim
    out_r = out.sreshape(ones) if A.ndim else out
    if init is not None:
        _fill_scalar(out, init)  # out[ParentIndex(1)] = nt.init  (:68)
    elif opc == abi.SB_OP_ADD:  # _init_reduction!  (:182-187)
        _fill_scalar(out, 0)
    elif opc == abi.SB_OP_MUL:
        _fill_scalar(out, 1)
    else:  # min / max start from f(first(A))  (:62, :184-185)
        first = A[tuple(slice(0, 1) for _ in range(A.ndim))]
        run_mapreduce(tokens, abi.SB_OP_NONE, abi.SB_INIT_NONE, 0.0, ones, [out_r, first])
    _mapreducedim_(tokens, opc, None, A.size, (out_r, A))
    host = out.to_numpy()
    v = host.reshape(-1)[0]
    return v.item()


def mapreduce(f, op, A, dims=None, init=None):
    """Base.mapreduce(f, op, A::StridedView; dims=:, init)   (mapreduce.jl:16-30).  `dims` is 0-based."""
    A = maybestrided(A)
    if dims is None:
        return _mapreduce_all(f, op, A, init)
    if isinstance(dims, int):
        dims = (dims,)
    dims = tuple(int(d) for d in dims)
    tokens = trace(f, 1)
    opc = _op_code(op)
    outsize = tuple(1 if d in dims else s for d, s in enumerate(A.size))
    out = similar_parent(A, result_dtype(tokens, [A]), outsize)
    # Base.reducedim_init / reducedim_initarray
    if init is not None:
        _fill_scalar(out, init)
    elif opc == abi.SB_OP_ADD:
        _fill_scalar(out, 0)
    elif opc == abi.SB_OP_MUL:
        _fill_scalar(out, 1)
    else:  # min/max: initialise with f of the first slice along the reduced dims
        if len(A) == 0:
            raise ValueError("reducing over an empty collection is not allowed")
        first = A[tuple(slice(0, 1) if d in dims else slice(None) for d in range(A.ndim))]
        run_mapreduce(tokens, abi.SB_OP_NONE, abi.SB_INIT_NONE, 0.0, outsize, [out, first])
    return mapreducedim_(tokens, opc, out, A)


def sum(A, dims=None):  # noqa: A001
    return mapreduce("identity", "+", A, dims=dims)


def prod(A, dims=None, f="identity"):
    return mapreduce(f, "*", A, dims=dims)


def maximum(A, dims=None, f="identity"):
    return mapreduce(f, "max", A, dims=dims)


def minimum(A, dims=None, f="identity"):
    return mapreduce(f, "min", A, dims=dims)


# ---- BLAS-1 style wrappers: one-line broadcasts  (linalg.jl:2-42) -------------------------------------------
def rmul_(dst, alpha):
    dst = maybestrided(dst)
    return materialize_(dst, Broadcasted("mul", (dst, alpha)))


def lmul_(alpha, dst):
    dst = maybestrided(dst)
    return materialize_(dst, Broadcasted("mul", (alpha, dst)))


def mul_(dst, a, b):
    """mul!(dst, alpha::Number, src) / mul!(dst, src, alpha::Number)   (linalg.jl:5-22)."""
    dst = maybestrided(dst)
    return materialize_(dst, Broadcasted("mul", (maybestrided(a), maybestrided(b))))


def axpy_(a, X, Y):
    """Y .= a .* X .+ Y   (linalg.jl:23-31)."""
    X, Y = maybestrided(X), maybestrided(Y)
    return materialize_(Y, Broadcasted("add", (Broadcasted("mul", (a, X)), Y)))


def axpby_(a, X, b, Y):
    """Y .= a .* X .+ b .* Y   (linalg.jl:32-42)."""
    X, Y = maybestrided(X), maybestrided(Y)
    return materialize_(Y, Broadcasted("add", (Broadcasted("mul", (a, X)), Broadcasted("mul", (b, Y)))))


# ---- generic matrix multiplication as a 3-D map-reduce  (linalg.jl:44-49, 130-162) ---------------------------
def _mul_generic_call(C, A, B, alpha=1, beta=0):
    """The `_mapreducedim!` call `__mul!(C, A, B, α, β)` makes (linalg.jl:130-162):  C2 = sreshape(C, (m,n,1)),
    A2 = sreshape(A, (m,1,k)), B2 = sreshape(permutedims(B,(2,1)), (1,n,k)), f = * or (x,y)->x*y*α, op = +,
    initop = zero | nothing | x->x*β.  Returns None when the reference short-cuts to `rmul!(C, β)` (α == 0 or k == 0),
    else (f, op, initop, dims, arrays)."""
    C, A, B = maybestrided(C), maybestrided(A), maybestrided(B)
    if not (C.ndim == A.ndim == B.ndim == 2 and C.size[0] == A.size[0] and C.size[1] == B.size[1] and A.size[1] == B.size[0]):
        raise DimensionMismatchError(f"A has size {A.size}, B has size {B.size}, C has size {C.size}")  # linalg.jl:132-133
    (m, n), k = C.size, A.size[1]
    if alpha == 0 or k == 0:
        return None
    A2 = A.sreshape((m, 1, k))
    B2 = B.permutedims((1, 0)).sreshape((1, n, k))
    C2 = C.sreshape((m, n, 1))
    f = "mul" if alpha == 1 else (lambda x, y: x * y * alpha)
    initop = "zero" if beta == 0 else (None if beta == 1 else ("scale", beta))
    return f, "+", initop, (m, n, k), (C2, A2, B2)


def mul_generic_(C, A, B, alpha=1, beta=0):
    """`mul!(C, A, B, α, β)` through the generic path `__mul!` (linalg.jl:44-49, 130-162): one fused
    initop + map + reduce launch, no temporaries.  (For BLAS eltypes the reference dispatches to `gemm!`,
    linalg.jl:50-127 -- a dense contraction, out of this engine's scope: call cuBLAS for that.)"""
    call = _mul_generic_call(C, A, B, alpha, beta)
    if call is None:
        return rmul_(C, beta)
    f, op, initop, dims, arrays = call
    _mapreducedim_(f, op, initop, dims, arrays)
    return maybestrided(C)
