"""CPU suite: random element programs (postfix trees over every function id of the ABI, stack depth <= 4) on dense
operands, through (a) the in-kernel interpreter as compiled for the host by the thread-grid emulator and (b) the C
restatement's evaluator, against the NumPy semantic oracle.  Pins the function table (`sb_fn`, elem.hpp call1/call2,
oracle/ref_eval.inc) across the three implementations for real and complex element types."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from helpers import Case, ViewSpec, randn

UNARY = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13]
BINARY = [32, 33, 34, 35, 36, 37]  # (lt yields exact zeros: 1/(0+0im) is where NumPy, C99 and Julia legitimately differ)


def _well_defined(toks, args):
    """every intermediate value finite and no 0/0, x/0: evaluated with NumPy under errstate(raise)"""
    from oracle import semantic
    st_ = []
    try:
        with np.errstate(divide="raise", invalid="raise", over="raise"):
            for kind, a, re, im in toks:
                if kind == 0:
                    st_.append(args[a])
                elif kind == 1:
                    st_.append(np.full_like(args[0], re))
                elif a < 32:
                    st_.append(semantic._FN1[a](st_.pop()))
                else:
                    y = st_.pop()
                    x = st_.pop()
                    if a == 35 and np.any(y == 0):
                        return False
                    st_.append(semantic._FN2[a](x, y))
                if not np.all(np.isfinite(st_[-1])):
                    return False
    except FloatingPointError:
        return False
    return True


@st.composite
def programs(draw):
    dt = draw(st.sampled_from((np.float32, np.float64, np.complex64, np.complex128)))
    nin = draw(st.integers(1, 3))
    toks, depth = [], 0
    target = draw(st.integers(1, 9))
    used = set()
    for _ in range(60):
        can_push = depth < 4
        can_un = depth >= 1
        can_bin = depth >= 2
        choices = (["push"] * 2 if can_push else []) + (["un"] if can_un else []) + (["bin"] * 2 if can_bin else [])
        if len(toks) >= target and depth == 1 and len(used) == nin:
            break
        if len(toks) >= target:  # wind down
            choices = ["bin"] if can_bin else (["push"] if len(used) < nin and can_push else ["un"])
        c = draw(st.sampled_from(choices))
        if c == "push":
            if draw(st.integers(0, 3)) == 0 and len(used) == nin:
                toks.append((1, 0, float(draw(st.sampled_from([2.0, 0.5, -1.5, 3.0]))), 0.0))
            else:
                missing = [k for k in range(nin) if k not in used]
                k = missing[0] if missing else draw(st.integers(0, nin - 1))
                used.add(k)
                toks.append((0, k, 0.0, 0.0))
            depth += 1
        elif c == "un":
            toks.append((2, draw(st.sampled_from(UNARY)), 0.0, 0.0))
        else:
            toks.append((2, draw(st.sampled_from(BINARY)), 0.0, 0.0))
            depth -= 1
    if depth != 1 or len(used) != nin or len(toks) > 40:
        toks = [(0, k, 0.0, 0.0) for k in range(nin)] + [(2, 32, 0.0, 0.0)] * (nin - 1)
    seed = draw(st.integers(0, 2 ** 31 - 1))
    return dt, nin, toks, seed


@settings(max_examples=300, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(programs())
def test_random_programs_agree_across_interpreter_restatement_and_numpy(spec):
    dt, nin, toks, seed = spec
    rng = np.random.default_rng(seed)
    n = 96
    # positive, moderate magnitudes: log / sqrt / division stay in their domains, exp does not overflow after a few levels
    parents = [np.zeros(n, dt)] + [(np.abs(randn(rng, n, dt)) * 0.5 + 0.25).astype(dt) for _ in range(nin)]
    views = [ViewSpec.dense(k, (n,)) for k in range(nin + 1)]
    tol = 3e-4 if dt in (np.float32, np.complex64) else 1e-10
    case = Case(f"prog_{np.dtype(dt).name}_{len(toks)}", parents, views, toks, rtol=tol)
    want = case.expected()
    if not np.all(np.isfinite(want)) or not _well_defined(toks, [p.astype(want.dtype if want.dtype.kind == "c" or dt in (np.float32, np.float64) else dt)
                                                                  for p in parents[1:]]):
        return  # overflow or a division by exact zero somewhere inside: NumPy, C99 and Julia legitimately differ there
    case.assert_close(case.run_emul(), want)
    case.assert_close(case.run_ref(1), want)
