"""TEST INFRASTRUCTURE ONLY: NumPy statement of the semantics of one `_mapreduce_fuse!` call.

Independent of any planner: views become `as_strided` windows on the same byte buffers, `f` is evaluated by
vectorised NumPy in the promoted compute type, reductions by `np.add.reduce`-style folds.  This restates what
the reference's tests use as THEIR oracle -- Base Julia on `Array`s (test/othertests.jl) -- not the
reference's implementation.
"""
from __future__ import annotations

import numpy as np

SB_TOK_ARG, SB_TOK_CONST, SB_TOK_CALL = 0, 1, 2
_FN1 = {
    0: lambda x: x, 1: lambda x: -x, 2: np.conj, 3: np.abs, 4: lambda x: (x.real * x.real + x.imag * x.imag) if np.iscomplexobj(x) else x * x,
    5: np.real, 6: np.imag, 7: np.sqrt, 8: np.exp, 9: np.log, 10: np.sin, 11: np.cos, 12: np.tanh, 13: lambda x: 1 / x,
}


def _jlmax(a, b):
    r = np.maximum(a, b)  # NaN-propagating like Julia's max
    return r


def _jlmin(a, b):
    return np.minimum(a, b)


_FN2 = {
    32: lambda x, y: x + y, 33: lambda x, y: x - y, 34: lambda x, y: x * y, 35: lambda x, y: x / y,
    36: lambda x, y: _jlmax(np.real(x), np.real(y)), 37: lambda x, y: _jlmin(np.real(x), np.real(y)),
    38: lambda x, y: (np.real(x) < np.real(y)).astype(np.result_type(x, y)),
}
_NPDT = {0: np.float32, 1: np.float64, 2: np.complex64, 3: np.complex128}


def compute_dtype(tokens, in_dtypes, out_dtype=None, reduce=False, init=0.0):
    cplx = any(d in (2, 3) for d in in_dtypes)
    dbl = any(d in (1, 3) for d in in_dtypes)
    for kind, a, re, im in tokens:
        if kind == SB_TOK_CONST:
            cplx |= im != 0.0
            dbl |= a == 2
    if reduce and out_dtype is not None:
        cplx |= out_dtype in (2, 3)
        dbl |= out_dtype in (1, 3)
        cplx |= complex(init).imag != 0.0
    return _NPDT[(2 if cplx else 0) + (1 if dbl else 0)]


def evaluate(tokens, args, ct):
    """args: list of ndarrays already broadcast to the full dims, in compute type `ct`."""
    if not tokens:
        return args[0]
    st = []
    with np.errstate(all="ignore"):
        for kind, a, re, im in tokens:
            if kind == SB_TOK_ARG:
                st.append(args[a])
            elif kind == SB_TOK_CONST:
                st.append(ct(complex(re, im)) if np.issubdtype(ct, np.complexfloating) else ct(re))
            elif a < 32:
                x = st.pop()
                r = _FN1[a](x)
                st.append(np.asarray(r).astype(ct, copy=False) if np.issubdtype(ct, np.complexfloating) or not np.iscomplexobj(r) else r)
            else:
                y = st.pop()
                x = st.pop()
                st.append(np.asarray(_FN2[a](x, y)).astype(ct, copy=False))
    return st[-1]


def as_window(flat, offset, size, strides):
    """flat: 1-D ndarray (the parent); element strides/offset as in StridedView."""
    its = flat.itemsize
    lo = offset + sum(min((n - 1) * s, 0) for n, s in zip(size, strides) if n > 0)
    base = flat[lo:]
    return np.lib.stride_tricks.as_strided(base[offset - lo:], shape=size, strides=tuple(s * its for s in strides), writeable=False)


def mapreduce(tokens, op, initop, init, dims, out_spec, in_specs):
    """Returns the new FULL-RANK output window values (shape: dims with 1 where the output stride is 0).

    *_spec = (flat_parent, offset, strides, dtype_code, conj).  The output parent is NOT modified."""
    dims = tuple(dims)
    oflat, ooff, ostr, odt, ocj = out_spec
    in_dt = [s[3] for s in in_specs]
    ct = compute_dtype(tokens, in_dt if in_specs else [odt], odt, reduce=op != 0, init=init)
    args = []
    for flat, off, strd, dt, cj in in_specs:
        w = as_window(flat, off, dims, strd)
        w = np.conj(w) if cj else w
        args.append(np.asarray(w).astype(ct))
    if not args:
        args = [np.zeros(dims, dtype=ct)]
    val = evaluate(tokens, args, ct)
    val = np.broadcast_to(np.asarray(val).astype(ct), dims)
    red_axes = tuple(i for i, (n, s) in enumerate(zip(dims, ostr)) if s == 0 and n != 1)
    outshape = tuple(1 if i in red_axes else n for i, n in enumerate(dims))
    ostr_k = tuple(0 if i in red_axes else s for i, s in enumerate(ostr))
    if op == 0:
        res = val
    else:
        old = as_window(oflat, ooff, outshape, ostr_k)
        old = (np.conj(old) if ocj else old).astype(ct)
        beta = ct(complex(init)) if np.issubdtype(ct, np.complexfloating) else ct(complex(init).real)
        if initop == 1:
            old = np.zeros_like(old)
        elif initop == 3:
            old = beta * old
        elif initop == 4:
            old = np.full_like(old, beta)
        elif initop == 5:
            old = np.conj(old)
        with np.errstate(all="ignore"):
            if op == 1:
                part = val.sum(axis=red_axes, keepdims=True, dtype=ct) if red_axes else val
                res = old + part
            elif op == 2:
                part = val.prod(axis=red_axes, keepdims=True, dtype=ct) if red_axes else val
                res = old * part
            elif op == 3:
                part = np.real(val).min(axis=red_axes, keepdims=True) if red_axes else np.real(val)
                res = np.minimum(np.real(old), part)
            else:
                part = np.real(val).max(axis=red_axes, keepdims=True) if red_axes else np.real(val)
                res = np.maximum(np.real(old), part)
    # conversion to the output eltype (real part when the destination is real)
    npout = _NPDT[odt]
    if not np.issubdtype(npout, np.complexfloating):
        res = np.real(res)
    res = np.asarray(res).astype(npout)
    return np.conj(res) if ocj else res
