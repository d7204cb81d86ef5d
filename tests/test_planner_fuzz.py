"""CPU suite: the host planner (sb_plan_describe -> build_plan) must answer EVERY descriptor with a status code -- never
crash, hang or divide by zero -- including absurd extents (2^33 per dim), huge and negative strides, misaligned bases
and every operator / initop combination.  (A fuzz run of this kind found an int64 overflow of the tile-step product
that ended in a division by zero; index spaces beyond 2^48 elements are now rejected up front.)"""
import ctypes as C

import numpy as np

from helpers import sb


def _random_desc(rng):
    d = sb.abi.sb_desc()
    n, m = int(rng.integers(0, 9)), int(rng.integers(1, 9))
    d.ndim, d.nops = n, m
    big = rng.random() < 0.3
    for i in range(n):
        d.dims[i] = int(rng.choice([0, 1, 2, 3, 7, 64, 1000, 4096, 2 ** 20, 2 ** 31, 2 ** 33])) if big else int(rng.integers(0, 70))
    for k in range(m):
        for i in range(n):
            d.strides[k][i] = (int(rng.choice([0, 1, -1, 2, 64, 4096, -4096, 2 ** 30, -(2 ** 35), 2 ** 40])) if rng.random() < 0.5
                               else int(rng.integers(-50, 50)))
        d.base[k] = int(rng.choice([0x7f0000000000, 0x7f0000000008, 0x7f0000001000, 0x7f0000000004]))
        d.dtype[k], d.conj[k] = int(rng.integers(0, 4)), int(rng.integers(0, 2))
    nin = m - 1
    if nin == 0:
        prog = [(1, 0, 2.0, 0.0)]
    elif nin == 1:
        prog = [[], [(0, 0, 0, 0), (2, 4, 0, 0)], [(1, 0, 3.0, 0), (0, 0, 0, 0), (2, 34, 0, 0)]][int(rng.integers(0, 3))]
    else:
        prog = [(0, 0, 0, 0)] + sum([[(0, k, 0, 0), (2, 32, 0, 0)] for k in range(1, nin)], [])
    if rng.random() < 0.1:  # malformed programs must be status codes too
        prog = prog + [(2, int(rng.choice([33, 99])), 0, 0)]
    d.ntok = len(prog)
    for i, (kd, a, re, im) in enumerate(prog):
        d.prog[i].kind, d.prog[i].a, d.prog[i].re, d.prog[i].im = kd, a, re, im
    d.op, d.initop, d.init_re = int(rng.integers(0, 5)), int(rng.integers(0, 6)), 0.5
    return d


def test_planner_answers_every_descriptor_with_a_status_code():
    lib = sb.abi.load_library()
    rng = np.random.default_rng(20261017)
    buf = C.create_string_buffer(1 << 16)
    codes = {}
    for _ in range(2500):
        rc = lib.sb_plan_describe(None, C.byref(_random_desc(rng)), buf, len(buf))
        codes[rc] = codes.get(rc, 0) + 1
    assert set(codes) <= {sb.abi.SB_OK, sb.abi.SB_E_INVALID, sb.abi.SB_E_SHAPE, sb.abi.SB_E_UNSUPPORTED}, codes
    assert codes.get(sb.abi.SB_OK, 0) > 1000, codes


def test_index_space_overflow_is_rejected_not_divided_by_zero():
    lib = sb.abi.load_library()
    d = sb.abi.sb_desc()
    dims = [8589934592, 1, 1048576, 1, 4096, 1048576, 64]
    d.ndim, d.nops = len(dims), 2
    for i, s in enumerate(dims):
        d.dims[i] = s
        d.strides[0][i] = 0 if i else 1
        d.strides[1][i] = 1
    d.base[0], d.base[1] = 0x7f0000000000, 0x7f0000100000
    d.dtype[0] = d.dtype[1] = 1
    d.op = 1
    buf = C.create_string_buffer(4096)
    assert lib.sb_plan_describe(None, C.byref(d), buf, len(buf)) == sb.abi.SB_E_UNSUPPORTED
    assert b"2^48" in lib.sb_last_error(None)
