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
    os.environ["SB_JIT_SYNC"] = "1"  # (by default the NVRTC compile runs in the background while the interpreter serves)
    sb.get_engine(0).reload_env()  # the SB_* knobs are read at ctx creation / on reload, not on the launch path
    yield
    os.environ.pop("SB_JIT_SYNC", None)
    if old is None:
        os.environ.pop("SB_JIT_MIN_ELEMENTS", None)
    else:
        os.environ["SB_JIT_MIN_ELEMENTS"] = old
    sb.get_engine(0).reload_env()


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
        sb.get_engine(0).reload_env()
        interp = c.run_gpu("device")
        os.environ["SB_JIT_MIN_ELEMENTS"] = "0"
        sb.get_engine(0).reload_env()
        assert jit.tobytes() == interp.tobytes(), c.name


def test_deep_expression_tree_runs_as_specialised_kernel():
    """an operand stack deeper than the interpreter's 4 register slots (the reference has no limit, src/broadcast.jl:67-98)
    is legal: the plan is marked needs_jit and runs as straight-line NVRTC code, at any problem size"""
    import numpy as np
    from helpers import A, F, Case, ViewSpec
    rng = np.random.default_rng(7)
    n = 3000
    xs = [rng.standard_normal(n) for _ in range(6)]
    # x0 + (x1 * (x2 - (x3 / (x4 + x5)))): postfix pushes all six operands before the first call -> depth 6
    toks = [A(0), A(1), A(2), A(3), A(4), A(5), F("add"), F("div"), F("sub"), F("mul"), F("add")]
    c = Case("deep6", [np.zeros(n)] + xs, [ViewSpec.dense(0, (n,))] + [ViewSpec.dense(k + 1, (n,)) for k in range(6)], toks)
    p = c.plan()
    assert p.get("needs_jit") == 1
    eng = sb.get_engine(0)
    eng.reset_stats()
    got = c.run_gpu("device")
    assert eng.stats()["jit_launches"] == 1
    want = xs[0] + (xs[1] * (xs[2] - (xs[3] / (xs[4] + xs[5]))))
    assert np.array_equal(got, want)


def test_background_compile_takes_over():
    """default mode: the first call of a new expression is served by the interpreter while NVRTC compiles on a worker
    thread; a later call runs the specialised kernel; both give the same bits"""
    import time
    import numpy as np
    from helpers import A, F, K, Case, ViewSpec
    rng = np.random.default_rng(11)
    n = 1 << 16
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    toks = [A(0), K(0.37, typ=2), F("mul"), A(1), F("sin"), F("sub"), F("abs")]  # abs(0.37 x - sin(y)): no recipe
    c = Case("bg_jit", [np.zeros(n), x, y], [ViewSpec.dense(0, (n,)), ViewSpec.dense(1, (n,)), ViewSpec.dense(2, (n,))], toks)
    os.environ["SB_JIT_MIN_ELEMENTS"] = "0"
    os.environ.pop("SB_JIT_SYNC", None)
    eng = sb.get_engine(0)
    eng.reload_env()
    try:
        eng.reset_stats()
        first = c.run_gpu("device")
        j0 = eng.stats()["jit_launches"]
        deadline = time.time() + 120
        later = first
        while time.time() < deadline:
            later = c.run_gpu("device")
            if eng.stats()["jit_launches"] > j0:
                break
            time.sleep(0.05)
        assert eng.stats()["jit_launches"] > j0, "the specialised kernel never took over"
        assert first.tobytes() == later.tobytes()
        c.assert_close(later)
    finally:
        os.environ.pop("SB_JIT_MIN_ELEMENTS", None)
        eng.reload_env()
