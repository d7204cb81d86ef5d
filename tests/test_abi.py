"""CPU suite: the C-ABI library loads, exports every symbol include/strided_b200.h declares, refuses to compute
without a GPU (no CPU fallback), and plans the BASELINE configs as designed (sb_plan_describe, host only)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import ROOT, case_c1, case_c2, case_c3, case_c4, case_c5, sb


def _declared():
    hdr = open(os.path.join(ROOT, "include", "strided_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", hdr)))


def test_exports_match_header():
    lib = sb.abi.load_library()
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/strided_b200.h but not exported"
    assert sorted(sb.abi.EXPORTS) == names
    assert lib.sb_abi_version() == 1


def test_struct_layout_matches_header():
    # sizeof(sb_desc) as laid out by the C compiler (natural alignment): checked against ctypes
    assert C.sizeof(sb.abi.sb_tok) == 24
    expect = 4 + 4 + 8 * 8 + 8 * 8 * 8 + 8 * 8 + 4 * 8 + 4 * 8 + 4 + 4 + 24 * 48 + 4 + 4 + 8 + 8
    assert C.sizeof(sb.abi.sb_desc) == expect


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = sb.abi.load_library()
    ctx = C.c_void_p()
    assert lib.sb_ctx_create(0, None, C.byref(ctx)) == sb.abi.SB_E_NODEVICE
    c = case_c1(16)
    with pytest.raises(sb.NoDeviceError):
        c.run_gpu("host")


def test_plans_for_baseline_configs():
    p = case_c2(4000).plan()
    assert p["family"] == "map_tile" and p["recipe"] == "add2_mul" and p["ct"] == "f64"
    assert p["tile"] == [64, 32] and p["staged"] == [0, 0, 1] and p["ntiles"] == 63 * 125
    assert p["tile_order"] == 1  # A and A' alias: tiles (I,J),(J,I) are launched side by side
    assert p["tma"] >= 2  # the TMA ring kernel
    # single-input transposes: small problems and tilings without edge tiles take the LSU kernel with per-tile records
    # (faster than the TMA ring there, profiles/r02_v_tma_vs_lsu_with_records.txt); edge tiles keep the TMA ring
    p = case_c1(1000).plan()
    assert p["recipe"] == "scale" and p["staged"] == [0, 1] and p["tile_order"] == 0
    assert p["tma"] == 0 and p["ept"] == 4 and p["tile"] == [32, 32] and p["lsu_desc"] == 1
    p = case_c3(32).plan()
    assert p["recipe"] == "copy" and p["dims"] == [32, 32, 32, 32] and p["tile"] == [32, 1, 1, 32] and p["ept"] == 4 and p["tma"] == 0 and p["lsu_desc"] == 1
    p = case_c3(128).plan()  # large, no edge tiles: LSU
    assert p["tma"] == 0 and p["lsu_desc"] == 1 and p["ept"] == 8
    p = case_c3(54).plan()   # 54 = 0.84 of a 64-wide tile: the TMA unit clips edge boxes for free
    assert p["tma"] >= 2
    p = case_c1(3000).plan()
    assert p["tma"] >= 2
    p = case_c5(1, 4096).plan()  # per-GPU share of config 5: one dense run -> streamed complete reduction
    assert p["dims"] == [16777216] and p["stream"]["grid"] == 148 and p["stream"]["chunk_bytes"] == 32768 and p["stream"]["nstage"] == 4
    p = case_c4(64).plan()
    assert p["recipe"] == "sum4" and p["ept"] == 16 and p["tile"] == [8, 8, 8, 8] and p["staged"] == [0, 0, 1, 1, 1]
    assert p["tile_order"] == 1
    p = case_c5(8, 4096).plan()
    assert p["family"] == "reduce_tile" and p["recipe"] == "abs2" and p["dims"] == [8, 16777216]
    assert p["tile"] == [8, 256] and p["nout_tile"] == 8 and p["nred_tile"] == 256 and p["nsplit"] > 100


def test_invalid_descriptors_are_status_codes():
    lib = sb.abi.load_library()
    buf = C.create_string_buffer(4096)
    d = case_c1(8).desc(case_c1(8).fresh())
    d.ndim = 99
    assert lib.sb_plan_describe(None, C.byref(d), buf, len(buf)) == sb.abi.SB_E_INVALID
    d = case_c1(8).desc(case_c1(8).fresh())
    d.dims[0] = -1
    assert lib.sb_plan_describe(None, C.byref(d), buf, len(buf)) == sb.abi.SB_E_SHAPE
    d = case_c1(8).desc(case_c1(8).fresh())
    d.strides[0][1] = 0  # map into a broadcast (zero-stride) destination: undefined in parallel
    assert lib.sb_plan_describe(None, C.byref(d), buf, len(buf)) == sb.abi.SB_E_UNSUPPORTED
    assert b"zero-stride" in lib.sb_last_error(None)


def test_bank_model_padding_is_conflict_free_for_transpose():
    # the staged operand of C2 must be readable/writable without shared-memory bank conflicts
    p = case_c2(512).plan()
    assert p["smem_bytes"] >= p["tile"][0] * p["tile"][1] * 8


def test_header_is_plain_c_and_layout_matches_ctypes(tmp_path):
    """include/strided_b200.h must compile as C99 (a Julia `ccall` / cgo / JNI binding sees C, not C++), and the
    struct sizes the C compiler computes must be the ones the ctypes mirror (and julia/StridedB200.jl) assume."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "strided_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%zu %zu %zu %d %d %d\\n", sizeof(sb_desc), sizeof(sb_tok), sizeof(sb_stats), '
                   'SB_PEER_MAX_OUT, SB_PEER_MAX_WORLD, SB_IPC_HANDLE_BYTES); return 0; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert int(out[0]) == C.sizeof(sb.abi.sb_desc) and int(out[1]) == C.sizeof(sb.abi.sb_tok)
    assert int(out[2]) == C.sizeof(sb.abi.sb_stats)
    assert (int(out[3]), int(out[4]), int(out[5])) == (sb.abi.SB_PEER_MAX_OUT, sb.abi.SB_PEER_MAX_WORLD, sb.abi.SB_IPC_HANDLE_BYTES)


def test_peer_entry_points_reject_bad_arguments_without_a_gpu():
    lib = sb.abi.load_library()
    buf = C.create_string_buffer(64)
    assert lib.sb_peer_export(None, buf) == sb.abi.SB_E_INVALID
    assert lib.sb_peer_attach(None, 0, 2, buf) == sb.abi.SB_E_INVALID
    assert lib.sb_peer_detach(None) == sb.abi.SB_OK
    d = case_c5(8, 16).desc(case_c5(8, 16).fresh())
    assert lib.sb_mapreduce_allreduce(None, C.byref(d)) == sb.abi.SB_E_INVALID
