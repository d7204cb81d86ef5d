"""Broadcasting over StridedViews: the Python mirror of reference src/broadcast.jl.

`Broadcasted(fn, args)` is the lazy fused tree Julia builds for a dotted expression.  `materialize_`
(= `copyto!(dest, bc)`, broadcast.jl:27-37) flattens it exactly like the reference:

  * `capturestridedargs` (:41-46): the StridedView leaves, depth-first, left to right -> operands 1..M-1;
  * `promoteshape` (:50-65): every view is promoted to dest's rank/size, size-1 dims get stride 0,
    anything else throws DimensionMismatch;
  * `make_capture` (:67-83): the tree with `Arg()` placeholders and baked-in scalars -> here a postfix
    token program (include/strided_b200.h `sb_tok`), arguments consumed in the same order as `consume`
    (:86-98);

and hands (program, dims, views) to the engine -- the ccall boundary that replaces `_mapreduce_fuse!`.
"""
from __future__ import annotations

import numbers

import numpy as np

from . import abi
from .view import StridedView, DimensionMismatchError, maybestrided

_UNARY = {"identity", "neg", "conj", "abs", "abs2", "real", "imag", "sqrt", "exp", "log", "sin", "cos", "tanh", "inv"}
_BINARY = {"add", "sub", "mul", "div", "max", "min", "lt"}


class Ref:
    """`Ref(x)`: a wrapped scalar, baked into the program (WrappedScalarArgs, broadcast.jl:39,81)."""

    def __init__(self, x):
        self.x = x


class Arg:
    """Placeholder for the i-th strided argument when tracing a plain callable `f` (map/mapreduce)."""

    def __init__(self, i):
        self.i = i

    def _bc(self, fn, *a):
        return Broadcasted(fn, a)

    __add__ = lambda s, o: s._bc("add", s, o)
    __radd__ = lambda s, o: s._bc("add", o, s)
    __sub__ = lambda s, o: s._bc("sub", s, o)
    __rsub__ = lambda s, o: s._bc("sub", o, s)
    __mul__ = lambda s, o: s._bc("mul", s, o)
    __rmul__ = lambda s, o: s._bc("mul", o, s)
    __truediv__ = lambda s, o: s._bc("div", s, o)
    __rtruediv__ = lambda s, o: s._bc("div", o, s)
    __neg__ = lambda s: s._bc("neg", s)
    __lt__ = lambda s, o: s._bc("lt", s, o)
    __gt__ = lambda s, o: s._bc("lt", o, s)


class Broadcasted:
    def __init__(self, fn, args):
        if fn not in _UNARY and fn not in _BINARY:
            raise abi.UnsupportedError(abi.SB_E_UNSUPPORTED, f"function {fn!r} is not in the device function set")
        self.fn = fn
        self.args = tuple(args)
        want = 1 if fn in _UNARY else 2
        if len(self.args) != want:
            raise TypeError(f"{fn} takes {want} argument(s)")

    def _bc(self, fn, *a):
        return Broadcasted(fn, a)

    __add__ = lambda s, o: s._bc("add", s, o)
    __radd__ = lambda s, o: s._bc("add", o, s)
    __sub__ = lambda s, o: s._bc("sub", s, o)
    __rsub__ = lambda s, o: s._bc("sub", o, s)
    __mul__ = lambda s, o: s._bc("mul", s, o)
    __rmul__ = lambda s, o: s._bc("mul", o, s)
    __truediv__ = lambda s, o: s._bc("div", s, o)
    __rtruediv__ = lambda s, o: s._bc("div", o, s)
    __neg__ = lambda s: s._bc("neg", s)
    __lt__ = lambda s, o: s._bc("lt", s, o)
    __gt__ = lambda s, o: s._bc("lt", o, s)

    def materialize(self):
        return materialize(self)


def _unary(name):
    def f(x):
        return Broadcasted(name, (x,))
    f.__name__ = name
    return f


def _binary(name):
    def f(x, y):
        return Broadcasted(name, (x, y))
    f.__name__ = name
    return f


# element functions usable inside broadcast expressions / traced lambdas
identity, neg, conj, abs_, abs2, real, imag, sqrt, exp, log, sin, cos, tanh, inv = (
    _unary(n) for n in ("identity", "neg", "conj", "abs", "abs2", "real", "imag", "sqrt", "exp", "log", "sin",
                        "cos", "tanh", "inv"))
add, sub, mul, div, maximum2, minimum2, lt = (_binary(n) for n in ("add", "sub", "mul", "div", "max", "min", "lt"))


def capturestridedargs(bc):
    """StridedView leaves, depth-first left-to-right (broadcast.jl:41-46)."""
    if isinstance(bc, StridedView):
        return [bc]
    if isinstance(bc, Broadcasted):
        out = []
        for a in bc.args:
            out.extend(capturestridedargs(a))
        return out
    return []


def promoteshape1(sz, a: StridedView):
    """broadcast.jl:56-65.  Julia aligns dims from the front: a rank-r view is padded with size-1 dims."""
    n = len(sz)
    if a.ndim > n:
        raise DimensionMismatchError("array could not be broadcasted to match destination")
    size = list(a.size) + [1] * (n - a.ndim)
    strides = list(a.strides) + [1] * (n - a.ndim)
    new = []
    for d in range(n):
        if size[d] == sz[d]:
            new.append(strides[d])
        elif size[d] == 1:
            new.append(0)
        else:
            raise DimensionMismatchError("array could not be broadcasted to match destination")
    return a._with(sz, new)


def promoteshape(sz, *views):
    return [promoteshape1(tuple(sz), v) for v in views]


def _const_token(x):
    """Scalar -> CONST token.  `a` tags the literal's type as Julia would see it: 0 = weak (Int/Rational/
    Bool: adopts the array eltype), 1 = Float32, 2 = Float64/ComplexF64."""
    if isinstance(x, Ref):
        x = x.x
    if isinstance(x, (bool, np.bool_)):
        return (abi.SB_TOK_CONST, 0, float(x), 0.0)
    if isinstance(x, (numbers.Integral, np.integer)):
        return (abi.SB_TOK_CONST, 0, float(int(x)), 0.0)
    if isinstance(x, np.floating):
        return (abi.SB_TOK_CONST, 2 if x.dtype == np.float64 else 1, float(x), 0.0)
    if isinstance(x, np.complexfloating):
        return (abi.SB_TOK_CONST, 2 if x.dtype == np.complex128 else 1, float(x.real), float(x.imag))
    if isinstance(x, float):
        return (abi.SB_TOK_CONST, 2, x, 0.0)
    if isinstance(x, complex):
        return (abi.SB_TOK_CONST, 2, x.real, x.imag)
    try:
        from fractions import Fraction
        if isinstance(x, Fraction):
            return (abi.SB_TOK_CONST, 0, float(x), 0.0)
    except Exception:  # pragma: no cover
        pass
    raise abi.UnsupportedError(abi.SB_E_UNSUPPORTED, f"cannot bake {type(x).__name__} into a device program")


def make_program(bc, counter=None):
    """make_capture + consume flattened to postfix (broadcast.jl:67-98).  The k-th StridedView leaf met in
    depth-first order becomes ARG k."""
    if counter is None:
        counter = [0]
    if isinstance(bc, StridedView):
        k = counter[0]
        counter[0] += 1
        return [(abi.SB_TOK_ARG, k, 0.0, 0.0)]
    if isinstance(bc, Arg):
        return [(abi.SB_TOK_ARG, bc.i, 0.0, 0.0)]
    if isinstance(bc, Broadcasted):
        toks = []
        for a in bc.args:
            toks.extend(make_program(a, counter))
        toks.append((abi.SB_TOK_CALL, abi.FN[bc.fn], 0.0, 0.0))
        return toks
    return [_const_token(bc)]


def trace(f, nargs):
    """Turn a callable (or a function name) into a postfix program over `nargs` arguments.  A Python lambda
    built from the functions of this module is traced symbolically; opaque callables cannot be introspected
    (neither could a precompiled engine introspect a Julia closure) -> UnsupportedError."""
    if f is None:
        return []
    if isinstance(f, str):
        if f not in abi.FN:
            raise abi.UnsupportedError(abi.SB_E_UNSUPPORTED, f"unknown function {f!r}")
        toks = [(abi.SB_TOK_ARG, i, 0.0, 0.0) for i in range(nargs)]
        want = 1 if f in _UNARY else 2
        if want != nargs:
            raise TypeError(f"{f} takes {want} argument(s), got {nargs}")
        return toks + [(abi.SB_TOK_CALL, abi.FN[f], 0.0, 0.0)]
    if isinstance(f, (list, tuple)):
        return list(f)
    try:
        res = f(*[Arg(i) for i in range(nargs)])
    except abi.UnsupportedError:
        raise
    except Exception as e:
        raise abi.UnsupportedError(abi.SB_E_UNSUPPORTED, f"callable is not traceable on the device path: {e}")
    return make_program(res)


def result_dtype(tokens, views):
    """Element type of `similar(bc, T)`: promotion over the arguments and typed constants."""
    cplx = any(v.is_complex for v in views)
    dbl = any(v.dtype in (abi.SB_F64, abi.SB_C64) for v in views)
    real_out = False
    for kind, a, re, im in tokens:
        if kind == abi.SB_TOK_CONST:
            cplx |= im != 0.0
            dbl |= a == 2
    # abs/abs2/real/imag/lt as the LAST call make the result real
    if tokens and tokens[-1][0] == abi.SB_TOK_CALL and abi.FN_NAME[tokens[-1][1]] in ("abs", "abs2", "real", "imag", "lt"):
        real_out = True
    if not views and not dbl:
        dbl = True
    return (2 if cplx and not real_out else 0) + (1 if dbl else 0)


def materialize_(dest, bc):
    """copyto!(dest::StridedView, bc::Broadcasted{StridedArrayStyle})   (broadcast.jl:27-37)."""
    from .engine import run_mapreduce
    dest = maybestrided(dest)
    if isinstance(bc, (StridedView, Arg)) or not isinstance(bc, Broadcasted):
        bc = Broadcasted("identity", (bc,))
    views = promoteshape(dest.size, *capturestridedargs(bc))
    tokens = make_program(bc)
    if len(dest) == 0:
        return dest
    run_mapreduce(tokens, abi.SB_OP_NONE, abi.SB_INIT_NONE, 0.0, dest.size, [dest] + views)
    return dest


def broadcast_shape(views):
    n = max((v.ndim for v in views), default=0)
    shape = [1] * n
    for v in views:
        for d, s in enumerate(v.size):
            if s != 1:
                if shape[d] != 1 and shape[d] != s:
                    raise DimensionMismatchError("arrays could not be broadcast to a common size")
                shape[d] = s
    return tuple(shape)


def materialize(bc):
    """`similar(bc, T)` + copyto!  (broadcast.jl:20-22, 27-37): allocate the result on the arguments' device."""
    from .engine import similar_parent
    views = capturestridedargs(bc)
    if not views:
        raise TypeError("broadcast expression contains no StridedView")
    tokens = make_program(bc)
    shape = broadcast_shape(views)
    dest = similar_parent(views[0], result_dtype(tokens, views), shape)
    return materialize_(dest, bc)
