"""The ccall boundary: builds an `sb_desc` from StridedViews and calls the C ABI.

This is what the Julia glue's method for device-backed views does in place of `_mapreduce_block!`
(reference src/mapreduce.jl:142): hand dims, per-operand element strides, base pointers, dtypes, conj
flags and the encoded (f, op, initop) to `sb_mapreduce`.  Device memory, streams and process groups come
from PyTorch (plumbing); every element is moved and computed by libstrided_b200.so.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import abi
from .view import StridedView, sb_to_numpy_dtype, sb_to_torch_dtype

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

_lock = threading.Lock()
_engines = {}


class Engine:
    """One sb_ctx (device + stream).  GPU analog of Strided.jl's thread-count globals (src/Strided.jl:18-35)."""

    def __init__(self, device=0):
        self.lib = abi.load_library()
        self.device = int(device)
        ctx = C.c_void_p()
        rc = self.lib.sb_ctx_create(self.device, None, C.byref(ctx))
        abi.check(self.lib, None, rc)
        self.ctx = ctx
        self._stream = None
        self.sync = True
        self.peer_world = 1

    def close(self):
        if self.ctx:
            self.lib.sb_ctx_destroy(self.ctx)
            self.ctx = None

    def set_stream(self, stream_ptr):
        if stream_ptr != self._stream:
            abi.check(self.lib, self.ctx, self.lib.sb_ctx_set_stream(self.ctx, C.c_void_p(stream_ptr)))
            self._stream = stream_ptr

    def set_sync(self, sync: bool):
        abi.check(self.lib, self.ctx, self.lib.sb_ctx_set_sync(self.ctx, 1 if sync else 0))
        self.sync = bool(sync)

    def reload_env(self):
        """SB_* knobs are read at ctx creation; call this after changing them (tests, tuning tools)."""
        abi.check(self.lib, self.ctx, self.lib.sb_ctx_reload_env(self.ctx))

    def synchronize(self):
        abi.check(self.lib, self.ctx, self.lib.sb_sync(self.ctx))

    def stats(self):
        st = abi.sb_stats()
        abi.check(self.lib, self.ctx, self.lib.sb_get_stats(self.ctx, C.byref(st)))
        return {k: int(getattr(st, k)) for k, _ in abi.sb_stats._fields_}

    def reset_stats(self):
        abi.check(self.lib, self.ctx, self.lib.sb_reset_stats(self.ctx))

    # ---- peer group (reductions across GPUs through peer memory, include/strided_b200.h "sb_peer_*") ----
    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(abi.SB_IPC_HANDLE_BYTES)
        abi.check(self.lib, self.ctx, self.lib.sb_peer_export(self.ctx, buf))
        return buf.raw

    def peer_attach(self, rank: int, world: int, handles):
        blob = b"".join(handles)
        assert len(blob) == world * abi.SB_IPC_HANDLE_BYTES
        abi.check(self.lib, self.ctx, self.lib.sb_peer_attach(self.ctx, int(rank), int(world), blob))
        self.peer_world = int(world)

    def peer_detach(self):
        abi.check(self.lib, self.ctx, self.lib.sb_peer_detach(self.ctx))
        self.peer_world = 1

    def mapreduce_allreduce(self, desc: abi.sb_desc):
        abi.check(self.lib, self.ctx, self.lib.sb_mapreduce_allreduce(self.ctx, C.byref(desc)))

    def mapreduce_batch(self, descs):
        """sb_mapreduce_batch: independent map calls overlap on side streams (device pointers)"""
        arr = (abi.sb_desc * len(descs))(*descs)
        abi.check(self.lib, self.ctx, self.lib.sb_mapreduce_batch(self.ctx, len(descs), arr))

    def mapreduce(self, desc: abi.sb_desc, host: bool):
        fn = self.lib.sb_mapreduce_host if host else self.lib.sb_mapreduce
        abi.check(self.lib, self.ctx, fn(self.ctx, C.byref(desc)))


def get_engine(device=0) -> Engine:
    with _lock:
        e = _engines.get(device)
        if e is None:
            e = Engine(device)
            _engines[device] = e
        return e


def make_desc(tokens, op, initop, init, dims, views) -> abi.sb_desc:
    d = abi.sb_desc()
    n, m = len(dims), len(views)
    if n > abi.SB_MAX_DIMS or m > abi.SB_MAX_OPS or len(tokens) > abi.SB_MAX_TOKENS:
        raise abi.UnsupportedError(abi.SB_E_UNSUPPORTED, "rank, operand count or program length beyond the ABI limits")
    d.ndim, d.nops = n, m
    for i, s in enumerate(dims):
        d.dims[i] = int(s)
    for k, v in enumerate(views):
        if v.ndim != n:
            raise ValueError("all operands must have the rank of dims (promoteshape them first)")
        for i in range(n):
            d.strides[k][i] = v.strides[i]
        d.base[k] = v.base_ptr()
        d.dtype[k] = v.dtype
        d.conj[k] = 1 if v.conj_flag else 0
    d.ntok = len(tokens)
    for i, (kind, a, re, im) in enumerate(tokens):
        d.prog[i].kind, d.prog[i].a, d.prog[i].re, d.prog[i].im = kind, a, re, im
    d.op, d.initop = op, initop
    init = complex(init)
    d.init_re, d.init_im = init.real, init.imag
    return d


def run_mapreduce(tokens, op, initop, init, dims, views, engine=None, allreduce=False):
    """_mapreduce_fuse!(f, op, initop, dims, arrays) on the device (reference src/mapreduce.jl:98-99).

    `engine`: an explicit :class:`Engine` (its own ctx + stream), e.g. two of them to overlap the H2D of one call
    with the D2H of the previous one on host-resident operands; default: the per-device engine."""
    desc = make_desc(tokens, op, initop, init, dims, views)
    devs = {v.device for v in views}
    if len(devs) != 1:
        raise ValueError(f"operands live on different devices: {sorted(devs)}")
    dev = devs.pop()
    if dev.startswith("cuda"):
        idx = int(dev.split(":")[1]) if ":" in dev else torch.cuda.current_device()
        eng = engine or get_engine(idx)
        if engine is None:
            eng.set_stream(torch.cuda.current_stream(idx).cuda_stream)
        if allreduce:  # collective over the engine's peer group (sb_mapreduce_allreduce)
            eng.mapreduce_allreduce(desc)
        else:
            eng.mapreduce(desc, host=False)
    else:  # host parents (plain `Array`s): staged through the device by sb_mapreduce_host
        if allreduce:
            raise ValueError("allreduce=True needs device-resident operands")
        eng = engine or get_engine(0)
        eng.mapreduce(desc, host=True)
    return views[0]


def run_batch(calls, engine=None):
    """A block of `_mapreduce_fuse!` calls as ONE batch: `calls` = [(tokens, op, initop, init, dims, views), ...] on device
    operands.  Independent map calls overlap (include/strided_b200.h: sb_mapreduce_batch); results are those of running the
    calls in order."""
    descs = [make_desc(*c) for c in calls]
    devs = {v.device for c in calls for v in c[5]}
    if len(devs) != 1 or not next(iter(devs)).startswith("cuda"):
        raise ValueError("run_batch needs device-resident operands on one GPU")
    dev = devs.pop()
    idx = int(dev.split(":")[1]) if ":" in dev else torch.cuda.current_device()
    eng = engine or get_engine(idx)
    if engine is None:
        eng.set_stream(torch.cuda.current_stream(idx).cuda_stream)
    eng.mapreduce_batch(descs)
    return [c[5][0] for c in calls]


def similar_parent(like: StridedView, dtype_code: int, shape):
    """`similar(a, T, dims)`: a fresh dense column-major parent on the same device, wrapped as a view."""
    n = 1
    for s in shape:
        n *= s
    strides, acc = [], 1
    for s in shape:
        strides.append(acc)
        acc *= s
    if like.is_device:
        flat = torch.empty(max(n, 1), dtype=sb_to_torch_dtype(dtype_code), device=like.device)
    else:
        flat = np.empty(max(n, 1), dtype=sb_to_numpy_dtype(dtype_code))
    return StridedView(flat, tuple(shape), tuple(strides), 0, False)
