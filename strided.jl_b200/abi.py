"""ctypes mirror of include/strided_b200.h (the C ABI) and the loader of libstrided_b200.so.

The library is the product; this module only binds it.  There is no Python/NumPy fallback: if the shared
library is missing, importing the compute entry points raises, and on a machine without a GPU
``sb_ctx_create`` returns SB_E_NODEVICE (surfaced as :class:`NoDeviceError`).
"""
from __future__ import annotations

import ctypes as C
import os

SB_MAX_DIMS = 8
SB_MAX_OPS = 8
SB_MAX_TOKENS = 48

SB_OK = 0
SB_E_INVALID, SB_E_SHAPE, SB_E_UNSUPPORTED, SB_E_CUDA, SB_E_NOMEM, SB_E_NODEVICE = -1, -2, -3, -4, -5, -6

SB_F32, SB_F64, SB_C32, SB_C64 = 0, 1, 2, 3
SB_OP_NONE, SB_OP_ADD, SB_OP_MUL, SB_OP_MIN, SB_OP_MAX = 0, 1, 2, 3, 4
SB_INIT_NONE, SB_INIT_ZERO, SB_INIT_IDENTITY, SB_INIT_SCALE, SB_INIT_CONST, SB_INIT_CONJ = 0, 1, 2, 3, 4, 5
SB_TOK_ARG, SB_TOK_CONST, SB_TOK_CALL = 0, 1, 2

FN = {
    "identity": 0, "neg": 1, "conj": 2, "abs": 3, "abs2": 4, "real": 5, "imag": 6, "sqrt": 7, "exp": 8,
    "log": 9, "sin": 10, "cos": 11, "tanh": 12, "inv": 13,
    "add": 32, "sub": 33, "mul": 34, "div": 35, "max": 36, "min": 37, "lt": 38,
}
FN_NAME = {v: k for k, v in FN.items()}


class sb_tok(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("re", C.c_double), ("im", C.c_double)]


class sb_desc(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32),
        ("nops", C.c_int32),
        ("dims", C.c_int64 * SB_MAX_DIMS),
        ("strides", (C.c_int64 * SB_MAX_DIMS) * SB_MAX_OPS),
        ("base", C.c_void_p * SB_MAX_OPS),
        ("dtype", C.c_int32 * SB_MAX_OPS),
        ("conj", C.c_int32 * SB_MAX_OPS),
        ("ntok", C.c_int32),
        ("prog", sb_tok * SB_MAX_TOKENS),
        ("op", C.c_int32),
        ("initop", C.c_int32),
        ("init_re", C.c_double),
        ("init_im", C.c_double),
    ]


class sb_stats(C.Structure):
    _fields_ = [("launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("plans_built", C.c_uint64), ("plans_cached", C.c_uint64), ("jit_launches", C.c_uint64),
                ("zero_copy_calls", C.c_uint64), ("batches", C.c_uint64), ("grouped_calls", C.c_uint64)]


class StridedB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"strided_b200 error {code}: {msg}")
        self.code = code


class DimensionMismatch(ValueError):
    """Same role as Julia's DimensionMismatch (reference src/mapreduce.jl:43-46, src/broadcast.jl:61)."""


class UnsupportedError(StridedB200Error):
    """SB_E_UNSUPPORTED: legal in the reference, not on the device path (the Julia glue falls back to CPU)."""


class NoDeviceError(StridedB200Error):
    pass


EXPORTS = [
    "sb_abi_version", "sb_ctx_create", "sb_ctx_destroy", "sb_ctx_set_stream", "sb_ctx_set_sync", "sb_ctx_reload_env", "sb_sync",
    "sb_last_error", "sb_malloc", "sb_free", "sb_memcpy_h2d", "sb_memcpy_d2h", "sb_mapreduce", "sb_mapreduce_batch",
    "sb_mapreduce_host", "sb_plan_describe", "sb_get_stats", "sb_reset_stats",
    "sb_peer_export", "sb_peer_attach", "sb_peer_detach", "sb_mapreduce_allreduce", "sb_shutdown",
]
SB_PEER_MAX_OUT, SB_PEER_MAX_WORLD, SB_IPC_HANDLE_BYTES = 1024, 8, 64

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# STRIDED_B200_LIB: explicit path of the shared library (the Julia glue honours the same variable); default: the in-tree build
LIB_PATH = os.environ.get("STRIDED_B200_LIB") or os.path.join(_PKG_DIR, "libstrided_b200.so")
_lib = None


def load_library():
    """dlopen the in-tree CUDA library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32 = C.c_void_p, C.c_int
    lib.sb_abi_version.restype = i32
    lib.sb_ctx_create.argtypes = [i32, vp, C.POINTER(vp)]
    lib.sb_ctx_destroy.argtypes = [vp]
    lib.sb_ctx_set_stream.argtypes = [vp, vp]
    lib.sb_ctx_set_sync.argtypes = [vp, i32]
    lib.sb_ctx_reload_env.argtypes = [vp]
    lib.sb_sync.argtypes = [vp]
    lib.sb_last_error.argtypes = [vp]
    lib.sb_last_error.restype = C.c_char_p
    lib.sb_malloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.sb_free.argtypes = [vp, vp]
    lib.sb_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    lib.sb_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    lib.sb_mapreduce.argtypes = [vp, C.POINTER(sb_desc)]
    lib.sb_mapreduce_host.argtypes = [vp, C.POINTER(sb_desc)]
    lib.sb_mapreduce_batch.argtypes = [vp, i32, C.POINTER(sb_desc)]
    lib.sb_plan_describe.argtypes = [vp, C.POINTER(sb_desc), C.c_char_p, C.c_size_t]
    lib.sb_get_stats.argtypes = [vp, C.POINTER(sb_stats)]
    lib.sb_reset_stats.argtypes = [vp]
    lib.sb_peer_export.argtypes = [vp, C.c_char_p]
    lib.sb_peer_attach.argtypes = [vp, i32, i32, C.c_char_p]
    lib.sb_peer_detach.argtypes = [vp]
    lib.sb_mapreduce_allreduce.argtypes = [vp, C.POINTER(sb_desc)]
    for n in EXPORTS:
        if n != "sb_last_error":
            getattr(lib, n).restype = i32
    lib.sb_shutdown.argtypes = []
    # the interpreter must not run its exit handlers while a background NVRTC compile is still in flight (jit.cu)
    import atexit
    atexit.register(lib.sb_shutdown)
    _lib = lib
    return lib


def check(lib, ctx, rc):
    if rc == SB_OK:
        return
    msg = lib.sb_last_error(ctx)
    msg = msg.decode() if msg else ""
    if rc == SB_E_SHAPE:
        raise DimensionMismatch(msg)
    if rc == SB_E_UNSUPPORTED:
        raise UnsupportedError(rc, msg)
    if rc == SB_E_NODEVICE:
        raise NoDeviceError(rc, msg)
    raise StridedB200Error(rc, msg)


def plan_describe(desc: sb_desc) -> dict:
    """Host-only: the plan the C++ planner picks for `desc` (no GPU needed)."""
    import json
    lib = load_library()
    buf = C.create_string_buffer(8192)
    rc = lib.sb_plan_describe(None, C.byref(desc), buf, len(buf))
    check(lib, None, rc)
    return json.loads(buf.value.decode())
