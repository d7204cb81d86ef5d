"""StridedView: Python mirror of the layout contract the hot path consumes.

The reference type lives in StridedViews.jl (external to the reference tree; used at src/broadcast.jl:64,
src/macros.jl:36-38, src/mapreduce.jl:136,268,276-278).  A view is pure metadata over a dense parent:

    element (i_1..i_N)  (0-based here)  lives at  parent[offset + sum_d i_d * strides_d]

with strides in elements, any sign, possibly 0, and an element-wise `op` in {identity, conj}.  All view
transformations below are lazy and never touch data ("never actually operates on the data", README.md:164).
Python differences, stated once: dims and permutations are 0-based; `*` on views is the broadcast `.*`.
"""
from __future__ import annotations

import numpy as np

from . import abi

try:  # torch is plumbing for device memory only
    import torch
except Exception:  # pragma: no cover
    torch = None

_NP2SB = {np.dtype("float32"): abi.SB_F32, np.dtype("float64"): abi.SB_F64,
          np.dtype("complex64"): abi.SB_C32, np.dtype("complex128"): abi.SB_C64}
_SB2NP = {v: k for k, v in _NP2SB.items()}
_SB_SIZE = {abi.SB_F32: 4, abi.SB_F64: 8, abi.SB_C32: 8, abi.SB_C64: 16}


def _torch_dtype_code(dt):
    table = {torch.float32: abi.SB_F32, torch.float64: abi.SB_F64, torch.complex64: abi.SB_C32,
             torch.complex128: abi.SB_C64}
    if dt not in table:
        raise TypeError(f"unsupported eltype {dt}: the device path covers Float32/64 and ComplexF32/64")
    return table[dt]


def sb_to_torch_dtype(code):
    return {abi.SB_F32: torch.float32, abi.SB_F64: torch.float64, abi.SB_C32: torch.complex64,
            abi.SB_C64: torch.complex128}[code]


def sb_to_numpy_dtype(code):
    return _SB2NP[code]


class StridedView:
    """StridedView(parent[, size, strides, offset, conj]) -- the 5-argument constructor mirrors the one the
    reference calls at src/broadcast.jl:64."""

    __array_priority__ = 1000

    def __init__(self, parent, size=None, strides=None, offset=0, conj=False):
        if isinstance(parent, StridedView):
            p = parent
            self.parent, self.ptr0, self.dtype, self.device = p.parent, p.ptr0, p.dtype, p.device
            size = p.size if size is None else size
            strides = p.strides if strides is None else strides
            offset = p.offset + offset
            conj = p.conj_flag != bool(conj)
        elif torch is not None and isinstance(parent, torch.Tensor):
            self.parent = parent
            self.ptr0 = parent.data_ptr()
            self.dtype = _torch_dtype_code(parent.dtype)
            self.device = str(parent.device)
            if size is None:
                size, strides = tuple(parent.shape), tuple(parent.stride())
        elif isinstance(parent, np.ndarray):
            if parent.dtype not in _NP2SB:
                raise TypeError(f"unsupported eltype {parent.dtype}")
            self.parent = parent
            self.ptr0 = parent.ctypes.data
            self.dtype = _NP2SB[parent.dtype]
            self.device = "cpu"
            if size is None:
                its = parent.itemsize
                if any(s % its for s in parent.strides):
                    raise ValueError("parent strides are not a multiple of the element size")
                size, strides = tuple(parent.shape), tuple(s // its for s in parent.strides)
        else:
            raise TypeError("parent must be a torch.Tensor, a numpy.ndarray or a StridedView")
        self.size = tuple(int(s) for s in size)
        self.strides = tuple(int(s) for s in strides)
        if len(self.size) != len(self.strides):
            raise ValueError("size and strides must have the same length")
        self.offset = int(offset)
        self.conj_flag = bool(conj) and self.dtype in (abi.SB_C32, abi.SB_C64)

    # ---- basic queries ------------------------------------------------------------------------------
    @property
    def ndim(self):
        return len(self.size)

    @property
    def shape(self):
        return self.size

    @property
    def itemsize(self):
        return _SB_SIZE[self.dtype]

    @property
    def is_complex(self):
        return self.dtype in (abi.SB_C32, abi.SB_C64)

    @property
    def is_device(self):
        return self.device.startswith("cuda")

    def __len__(self):
        n = 1
        for s in self.size:
            n *= s
        return n

    length = __len__

    def base_ptr(self):
        """pointer(parent, offset+1) in the reference's words (src/mapreduce.jl:268)."""
        return self.ptr0 + self.offset * self.itemsize

    def _with(self, size, strides, offset=None, conj=None):
        v = StridedView.__new__(StridedView)
        v.parent, v.ptr0, v.dtype, v.device = self.parent, self.ptr0, self.dtype, self.device
        v.size, v.strides = tuple(size), tuple(strides)
        v.offset = self.offset if offset is None else offset
        v.conj_flag = self.conj_flag if conj is None else (conj and self.is_complex)
        return v

    # ---- lazy transformations (StridedViews.jl) ------------------------------------------------------
    def permutedims(self, p):
        p = tuple(int(i) for i in p)
        if sorted(p) != list(range(self.ndim)):
            raise ValueError(f"{p} is not a permutation of 0..{self.ndim - 1}")
        return self._with([self.size[i] for i in p], [self.strides[i] for i in p])

    def transpose(self):
        if self.ndim != 2:
            raise ValueError("transpose is defined for 2-D views")
        return self.permutedims((1, 0))

    def conj(self):
        return self._with(self.size, self.strides, conj=not self.conj_flag)

    def adjoint(self):
        return self.transpose().conj()

    @property
    def T(self):
        return self.transpose()

    @property
    def H(self):
        return self.adjoint()

    def sreshape(self, newsize):
        newsize = tuple(int(s) for s in newsize)
        n_old, n_new = 1, 1
        for s in self.size:
            n_old *= s
        for s in newsize:
            n_new *= s
        if n_old != n_new:
            raise DimensionMismatchError(f"cannot sreshape {self.size} to {newsize}")
        return self._with(newsize, _reshape_strides(newsize, self.size, self.strides))

    def __getitem__(self, idx):
        """sview: ints, slices (any step, incl. negative) and Ellipsis-free full indexing."""
        if not isinstance(idx, tuple):
            idx = (idx,)
        if len(idx) != self.ndim:
            raise IndexError("sview needs one index per dimension")
        size, strides, off = [], [], self.offset
        for i, (ix, n, s) in enumerate(zip(idx, self.size, self.strides)):
            if isinstance(ix, slice):
                start, stop, step = ix.indices(n)
                cnt = len(range(start, stop, step))
                off += start * s
                size.append(cnt)
                strides.append(s * step)
            else:
                k = int(ix)
                if k < 0:
                    k += n
                if not 0 <= k < n:
                    raise IndexError(f"index {ix} out of range in dim {i}")
                off += k * s
        return self._with(size, strides, offset=off)

    # ---- data access (tests / small cases) ---------------------------------------------------------------
    def to_numpy(self):
        """Materialise the view on the host as a dense column-major ndarray (Array(::StridedView),
        reference src/convert.jl:1-17 -- here by plain NumPy indexing, for tests and examples)."""
        flat = self._flat_numpy()
        its = flat.itemsize
        v = np.lib.stride_tricks.as_strided(flat[self.offset - self._flat_origin:], shape=self.size,
                                            strides=tuple(s * its for s in self.strides), writeable=False)
        out = np.array(v, order="F")
        return np.conj(out) if self.conj_flag else out

    def _flat_numpy(self):
        # flat host copy of the touched range [lo, hi] of the parent
        lo = hi = self.offset
        for n, s in zip(self.size, self.strides):
            if n == 0:
                self._flat_origin = self.offset
                return np.zeros(1, dtype=sb_to_numpy_dtype(self.dtype))
            ext = (n - 1) * s
            lo += min(ext, 0)
            hi += max(ext, 0)
        self._flat_origin = lo
        npdt = sb_to_numpy_dtype(self.dtype)
        if isinstance(self.parent, np.ndarray):
            import ctypes
            buf = (ctypes.c_char * ((hi - lo + 1) * self.itemsize)).from_address(self.ptr0 + lo * self.itemsize)
            return np.frombuffer(buf, dtype=npdt)
        flat = torch.empty(hi - lo + 1, dtype=self.parent.dtype, device=self.parent.device)
        src = torch.as_strided(self.parent, (hi - lo + 1,), (1,), self.parent.storage_offset() + lo)
        flat.copy_(src)
        return flat.cpu().numpy()

    # ---- broadcasting sugar: building lazy Broadcasted trees (reference src/broadcast.jl) -----------------
    def _bc(self, fn, *args):
        from .broadcast import Broadcasted
        return Broadcasted(fn, args)

    def __add__(self, o):
        return self._bc("add", self, o)

    def __radd__(self, o):
        return self._bc("add", o, self)

    def __sub__(self, o):
        return self._bc("sub", self, o)

    def __rsub__(self, o):
        return self._bc("sub", o, self)

    def __mul__(self, o):
        return self._bc("mul", self, o)

    def __rmul__(self, o):
        return self._bc("mul", o, self)

    def __truediv__(self, o):
        return self._bc("div", self, o)

    def __rtruediv__(self, o):
        return self._bc("div", o, self)

    def __neg__(self):
        return self._bc("neg", self)

    def assign(self, bc):
        """`dest .= bc`  ->  copyto!(dest, bc)   (reference src/broadcast.jl:27-37)."""
        from .broadcast import materialize_
        return materialize_(self, bc)

    def __repr__(self):
        return (f"StridedView(size={self.size}, strides={self.strides}, offset={self.offset}, "
                f"dtype={sb_to_numpy_dtype(self.dtype)}, conj={self.conj_flag}, device={self.device})")


class DimensionMismatchError(abi.DimensionMismatch):
    pass


def _reshape_strides(newsize, oldsize, oldstrides):
    """Strides of `sreshape`: succeed only when the new shape is expressible with strides (StridedViews.jl
    contract); otherwise raise, as the reference does."""
    newsize, oldsize, oldstrides = list(newsize), list(oldsize), list(oldstrides)
    out = []
    while newsize:
        d = newsize[0]
        if not oldsize:
            if any(x != 1 for x in newsize):
                raise DimensionMismatchError("sreshape: sizes do not match")
            out.extend([1] * len(newsize))
            return tuple(out)
        if d == oldsize[0]:
            out.append(oldstrides[0])
            newsize.pop(0), oldsize.pop(0), oldstrides.pop(0)
        elif d < oldsize[0]:
            if d == 0 or oldsize[0] % d:
                raise DimensionMismatchError("sreshape: new shape is not strided")
            out.append(oldstrides[0])
            oldsize[0] //= d
            oldstrides[0] *= d
            newsize.pop(0)
        else:  # d > oldsize[0]: the new dim must swallow the next old dim(s)
            if len(oldsize) < 2:
                raise DimensionMismatchError("sreshape: sizes do not match")
            if oldsize[0] == 1:
                oldsize.pop(0), oldstrides.pop(0)
            elif oldsize[1] == 1:
                oldsize.pop(1), oldstrides.pop(1)
            elif oldsize[0] * oldstrides[0] == oldstrides[1]:
                oldsize[1] *= oldsize[0]
                oldstrides[1] = oldstrides[0]
                oldsize.pop(0), oldstrides.pop(0)
            else:
                raise DimensionMismatchError("sreshape: new shape is not strided")
    if any(x != 1 for x in oldsize):
        raise DimensionMismatchError("sreshape: sizes do not match")
    return tuple(out)


def sreshape(a: StridedView, newsize):
    return a.sreshape(newsize)


def sview(a: StridedView, *idx):
    return a[idx]


def isstrided(a):
    return isinstance(a, StridedView)


def maybestrided(a):
    """`maybestrided` of the @strided macro (reference src/macros.jl:31-34): arrays become views."""
    if isinstance(a, StridedView):
        return a
    if (torch is not None and isinstance(a, torch.Tensor)) or isinstance(a, np.ndarray):
        return StridedView(a)
    return a
