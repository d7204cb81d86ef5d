"""TEST INFRASTRUCTURE ONLY: ctypes binding of the C restatement (oracle/strided_ref.c).  PARITY UNPINNED."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libstrided_ref.so")
REF_MAX_LEAVES = 256
_lib = None


def build(force=False):
    """Compile the restatement with gcc (recipe: oracle/Makefile)."""
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(_DIR, f)) > os.path.getmtime(LIB_PATH)
            for f in ("strided_ref.c", "ref_eval.inc", "strided_ref.h")):
        subprocess.check_call(["make", "-C", _DIR, "-s"])
    return LIB_PATH


def _abi():
    import strided_jl_b200 as sb  # only for the shared ctypes definition of sb_desc (the interface format)
    return sb.abi


class ref_plan_out(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32),
        ("fused_dims", C.c_int64 * 8),
        ("importance", C.c_int64 * 8),
        ("perm", C.c_int32 * 8),
        ("dims", C.c_int64 * 8),
        ("strides", (C.c_int64 * 8) * 8),
        ("costs", C.c_int64 * 8),
        ("blocks", C.c_int64 * 8),
        ("region_bytes", C.c_int64),
        ("complete_reduction", C.c_int32),
        ("nleaves", C.c_int32),
        ("leaf_dims", (C.c_int64 * 8) * REF_MAX_LEAVES),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB_PATH)
        abi = _abi()
        l.ref_plan.argtypes = [C.POINTER(abi.sb_desc), C.c_int, C.POINTER(ref_plan_out)]
        l.ref_plan.restype = C.c_int
        l.ref_mapreduce.argtypes = [C.POINTER(abi.sb_desc), C.c_int]
        l.ref_mapreduce.restype = C.c_int
        l.ref_last_error.restype = C.c_char_p
        _lib = l
    return _lib


def mapreduce(desc, nthreads=1):
    """Run the restated reference path on HOST pointers with `nthreads` tasks."""
    rc = lib().ref_mapreduce(C.byref(desc), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"ref_mapreduce failed ({rc}): {lib().ref_last_error().decode()}")


def plan(desc, nthreads=1):
    out = ref_plan_out()
    rc = lib().ref_plan(C.byref(desc), int(nthreads), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"ref_plan failed ({rc}): {lib().ref_last_error().decode()}")
    n = out.ndim
    return {
        "fused_dims": tuple(out.fused_dims[:n]), "importance": tuple(out.importance[:n]), "perm": tuple(out.perm[:n]),
        "dims": tuple(out.dims[:n]), "costs": tuple(out.costs[:n]), "blocks": tuple(out.blocks[:n]),
        "region_bytes": int(out.region_bytes), "complete_reduction": bool(out.complete_reduction),
        "nleaves": int(out.nleaves),
        "leaves": [tuple(out.leaf_dims[i][:n]) for i in range(min(out.nleaves, REF_MAX_LEAVES))],
        "strides": [tuple(out.strides[k][:n]) for k in range(desc.nops)],
    }
