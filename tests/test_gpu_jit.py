"""GPU suite: NVRTC-specialised element functions (csrc/jit.cu) against the oracles, and the interpreter they replace."""
import os

import numpy as np
import pytest

import cases
from helpers import sb

pytestmark = pytest.mark.gpu

PICK = ("lambda3_float32_N3", "lambda3_float64_N4", "lambda3_complex128_N2", "bc1_float64", "bc3_complex64", "bc2_float32",
        "sin_135_float64", "initop_scale_complex128", "count_neg_perm_float32", "prod_exp_float64", "dot_float64", "matmul_initop",
        "max_abs_dense_complex128", "real_into_complex", "mul_alpha_float64_N5")


@pytest.fixture
def jit_everything():
    old = os.environ.get("SB_JIT_MIN_ELEMENTS")
    os.environ["SB_JIT_MIN_ELEMENTS"] = "0"
    yield
    if old is None:
        os.environ.pop("SB_JIT_MIN_ELEMENTS", None)
    else:
        os.environ["SB_JIT_MIN_ELEMENTS"] = old


def test_jit_kernels_match_oracle(jit_everything):
    eng = sb.get_engine(0)
    todo = [c for c in cases.all_cases(0.3) if c.name in PICK]
    assert len(todo) >= 12
    eng.reset_stats()
    for c in todo:
        got = c.run_gpu("device")
        c.assert_close(got)
    st = eng.stats()
    assert st["jit_launches"] >= 10, st  # the interpreter cases really ran as run-time compiled kernels


def test_jit_equals_interpreter_bitwise(jit_everything):
    # same arithmetic, same order: the specialised kernel and the interpreter agree bit for bit
    todo = [c for c in cases.all_cases(0.3) if c.name in ("lambda3_float64_N4", "bc1_float64", "mul_alpha_float64_N5")]
    for c in todo:
        jit = c.run_gpu("device")
        os.environ["SB_JIT_MIN_ELEMENTS"] = str(1 << 62)
        interp = c.run_gpu("device")
        os.environ["SB_JIT_MIN_ELEMENTS"] = "0"
        assert jit.tobytes() == interp.tobytes(), c.name
