"""Shared test machinery: one `Case` = one `_mapreduce_fuse!` call on seeded data, runnable through
  * the NumPy semantic oracle            (oracle/semantic.py)
  * the C restatement of the reference   (oracle/strided_ref.c, any task count)
  * the CPU thread-grid emulation of the CUDA kernels (tests/emul/, index logic only)
  * the CUDA engine through the C ABI    (device pointers or host pointers)
Every runner returns the full output PARENT buffer, so untouched elements are checked too.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import strided_jl_b200 as sb  # noqa: E402


class _LazyOracle:
    """oracle/ is test infrastructure: it is imported only when a checker is actually invoked (tools/ scripts reuse the
    Case builders of this module but never touch the oracle)."""

    def __init__(self, name):
        self._name, self._mod = name, None

    def __getattr__(self, attr):
        if self._mod is None:
            import importlib
            self._mod = importlib.import_module(self._name)
        return getattr(self._mod, attr)


oref = _LazyOracle("oracle.ref")
semantic = _LazyOracle("oracle.semantic")

NPDT = {0: np.float32, 1: np.float64, 2: np.complex64, 3: np.complex128}
CODE = {np.dtype(v): k for k, v in NPDT.items()}
SEED = 1234  # mirrors Random.seed!(1234) of the reference's test/runtests.jl:7


def randn(rng, n, dt):
    dt = np.dtype(dt)
    if dt.kind == "c":
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(dt)
    return rng.standard_normal(n).astype(dt)


def rand(rng, n, dt):
    dt = np.dtype(dt)
    if dt.kind == "c":
        return (rng.random(n) + 1j * rng.random(n)).astype(dt)
    return rng.random(n).astype(dt)


def col_major_strides(shape):
    st, acc = [], 1
    for s in shape:
        st.append(acc)
        acc *= s
    return tuple(st)


class ViewSpec:
    """(parent index, offset, size, strides, conj) -- a StridedView over parents[parent]."""

    def __init__(self, parent, offset, size, strides, conj=False):
        self.parent, self.offset, self.size, self.strides, self.conj = parent, int(offset), tuple(size), tuple(strides), bool(conj)

    @staticmethod
    def dense(parent, shape):
        return ViewSpec(parent, 0, shape, col_major_strides(shape))

    def permutedims(self, p):
        return ViewSpec(self.parent, self.offset, [self.size[i] for i in p], [self.strides[i] for i in p], self.conj)

    def with_strides(self, size, strides):
        return ViewSpec(self.parent, self.offset, size, strides, self.conj)


class Case:
    def __init__(self, name, parents, views, tokens, op=0, initop=0, init=0.0, dims=None, rtol=None):
        """parents: list of flat ndarrays; views[0] is the output.  All views must already be promoted to `dims`."""
        self.name = name
        self.parents = [np.ascontiguousarray(p) for p in parents]
        self.views = views
        self.tokens = list(tokens)
        self.op, self.initop, self.init = op, initop, init
        self.dims = tuple(dims if dims is not None else views[0].size)
        self.rtol = rtol

    # ---- plumbing -----------------------------------------------------------------------------------
    def _svs(self, bufs):
        return [sb.StridedView(bufs[v.parent], self.dims, v.strides, v.offset, v.conj) for v in self.views]

    def desc(self, bufs):
        return sb.make_desc(self.tokens, self.op, self.initop, self.init, self.dims, self._svs(bufs))

    def fresh(self):
        return [p.copy() for p in self.parents]

    # ---- runners -------------------------------------------------------------------------------------
    def expected(self):
        """Output parent after the call, per the NumPy semantic oracle."""
        bufs = self.fresh()
        specs = [(bufs[v.parent], v.offset, v.strides, CODE[bufs[v.parent].dtype], v.conj) for v in self.views]
        res = semantic.mapreduce(self.tokens, self.op, self.initop, self.init, self.dims, specs[0], specs[1:])
        out = bufs[self.views[0].parent]
        v0 = self.views[0]
        ostr = tuple(0 if (s == 0 and n != 1) else s for n, s in zip(self.dims, v0.strides))
        if any(n == 0 for n in res.shape):
            return out
        lo = v0.offset + sum(min((n - 1) * s, 0) for n, s in zip(res.shape, ostr))
        w = np.lib.stride_tricks.as_strided(out[lo:][v0.offset - lo:], shape=res.shape,
                                            strides=tuple(s * out.itemsize for s in ostr), writeable=True)
        w[...] = res
        return out

    def run_ref(self, nthreads=1):
        bufs = self.fresh()
        d = self.desc(bufs)
        oref.mapreduce(d, nthreads)
        return bufs[self.views[0].parent]

    def run_emul(self, grid_limit=0):
        bufs = self.fresh()
        d = self.desc(bufs)
        lib = emul_lib()
        rc = lib.emul_mapreduce(C.byref(d), int(grid_limit))
        if rc != 0:
            raise RuntimeError(f"emul_mapreduce failed ({rc}): {lib.emul_last_error().decode()}")
        return bufs[self.views[0].parent]

    def run_gpu(self, mode="device", engine=None):
        import torch
        if mode == "host":
            bufs = self.fresh()
            sb.run_mapreduce(self.tokens, self.op, self.initop, self.init, self.dims, self._svs(bufs))
            return bufs[self.views[0].parent]
        dev = [torch.from_numpy(p.copy()).cuda() for p in self.parents]
        if engine is not None:
            engine.set_stream(torch.cuda.current_stream().cuda_stream)
        sb.run_mapreduce(self.tokens, self.op, self.initop, self.init, self.dims, self._svs(dev), engine=engine)
        torch.cuda.synchronize()
        return dev[self.views[0].parent].cpu().numpy()

    def plan(self):
        return sb.plan_describe(self.desc(self.fresh()))

    # ---- golden fixtures (tests/golden/*.npz) ---------------------------------------------------------
    def to_npz(self, expected, exact):
        d = {"name": np.array(self.name), "nparents": np.array(len(self.parents)), "dims": np.array(self.dims, dtype=np.int64),
             "tokens": np.array(self.tokens, dtype=np.float64).reshape(-1, 4), "op": np.array(self.op),
             "initop": np.array(self.initop), "init": np.array(complex(self.init)), "expected": expected,
             "exact": np.array(bool(exact)), "rtol": np.array(-1.0 if self.rtol is None else self.rtol),
             "vparent": np.array([v.parent for v in self.views]), "voffset": np.array([v.offset for v in self.views]),
             "vconj": np.array([v.conj for v in self.views]),
             "vstrides": np.array([list(v.strides) for v in self.views], dtype=np.int64).reshape(len(self.views), -1)}
        for i, p in enumerate(self.parents):
            d[f"parent{i}"] = p
        return d

    @staticmethod
    def from_npz(z):
        parents = [z[f"parent{i}"] for i in range(int(z["nparents"]))]
        dims = tuple(int(x) for x in z["dims"])
        views = [ViewSpec(int(p), int(o), dims, tuple(int(s) for s in st), bool(cj))
                 for p, o, st, cj in zip(z["vparent"], z["voffset"], z["vstrides"], z["vconj"])]
        tokens = [(int(t[0]), int(t[1]), float(t[2]), float(t[3])) for t in z["tokens"]]
        rtol = float(z["rtol"])
        init = complex(z["init"])
        return Case(str(z["name"]), parents, views, tokens, int(z["op"]), int(z["initop"]),
                    init if init.imag else init.real, dims, None if rtol < 0 else rtol)

    # ---- comparison ------------------------------------------------------------------------------------
    def tolerance(self):
        if self.rtol is not None:
            return self.rtol
        dt = self.parents[self.views[0].parent].dtype
        return 1e-5 if dt in (np.float32, np.complex64) else 1e-12

    def assert_close(self, got, want=None, exact=False):
        want = self.expected() if want is None else want
        assert got.shape == want.shape and got.dtype == want.dtype, (got.dtype, want.dtype)
        if exact:
            assert got.tobytes() == want.tobytes(), f"{self.name}: not bit-exact ({np.sum(got != want)} elements differ)"
            return
        scale = max(1.0, float(np.max(np.abs(want)))) if want.size else 1.0
        err = float(np.max(np.abs(got - want))) if want.size else 0.0
        assert np.array_equal(np.isnan(got), np.isnan(want)), f"{self.name}: NaN pattern differs"
        assert err <= self.tolerance() * scale * 16, f"{self.name}: max abs err {err:.3e} (scale {scale:.3e})"


_emul = None


def emul_lib():
    """Build (g++) and load tests/emul/libsb_emul.so -- the CPU emulation of the CUDA kernel bodies."""
    global _emul
    if _emul is not None:
        return _emul
    d = os.path.join(ROOT, "tests", "emul")
    so = os.path.join(d, "libsb_emul.so")
    csrc = os.path.join(ROOT, "strided.jl_b200", "csrc")
    srcs = [os.path.join(d, "emul.cpp"), os.path.join(csrc, "planner.cpp")]
    deps = srcs + [os.path.join(csrc, f) for f in ("common.hpp", "elem.hpp", "functors.hpp", "map_tile.hpp", "reduce_tile.hpp", "planner.hpp", "tma_tile.hpp", "orbit_tile.hpp", "reduce_stream.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-ffp-contract=off", "-o", so] + srcs)
    lib = C.CDLL(so)
    lib.emul_mapreduce.argtypes = [C.POINTER(sb.abi.sb_desc), C.c_int]
    lib.emul_mapreduce.restype = C.c_int
    lib.emul_last_error.restype = C.c_char_p
    _emul = lib
    return lib


# ---- program shorthands (postfix tokens) ----------------------------------------------------------------
A = lambda i: (0, i, 0.0, 0.0)  # noqa: E731
K = lambda re, im=0.0, typ=0: (1, typ, float(re), float(im))  # noqa: E731
F = lambda name: (2, sb.abi.FN[name], 0.0, 0.0)  # noqa: E731

P_COPY = []
P_SCALE3 = [K(3), A(0), F("mul")]                                   # C1: 3 .* A'
P_AVG = [A(0), A(1), F("add"), K(2), F("div")]                      # C2: (A .+ A') ./ 2
P_SUM4 = [A(0), A(1), F("add"), A(2), F("add"), A(3), F("add")]     # C4
P_ABS2 = [A(0), F("abs2")]                                          # C5
P_LAMBDA3 = [A(0), F("sin"), A(1), A(2), F("abs"), F("neg"), F("exp"), F("div"), F("add")]  # othertests.jl:36


# ---- the BASELINE configs (SURVEY.md section 8 d) as Cases, at any size ----------------------------------
def case_c1(n=1000, dt=np.float64, seed=SEED):
    rng = np.random.default_rng(seed)
    a, b = randn(rng, n * n, dt), np.zeros(n * n, dt)
    B, Av = ViewSpec.dense(0, (n, n)), ViewSpec.dense(1, (n, n)).permutedims((1, 0))
    return Case(f"C1_{n}", [b, a], [B, Av], P_SCALE3)


def case_c2(n=4000, dt=np.float64, seed=SEED):
    rng = np.random.default_rng(seed)
    a, b = randn(rng, n * n, dt), np.zeros(n * n, dt)
    Av = ViewSpec.dense(1, (n, n))
    return Case(f"C2_{n}", [b, a], [ViewSpec.dense(0, (n, n)), Av, Av.permutedims((1, 0))], P_AVG)


def case_c3(n=32, dt=np.float64, p=(3, 2, 1, 0), seed=SEED):
    rng = np.random.default_rng(seed)
    a, b = randn(rng, n ** 4, dt), np.zeros(n ** 4, dt)
    return Case(f"C3_{n}", [b, a], [ViewSpec.dense(0, (n,) * 4), ViewSpec.dense(1, (n,) * 4).permutedims(p)], P_COPY)


def case_c4(n=64, dt=np.float32, seed=SEED):
    rng = np.random.default_rng(seed)
    a, b = randn(rng, n ** 4, dt), np.zeros(n ** 4, dt)
    Av = ViewSpec.dense(1, (n,) * 4)
    perms = [(0, 1, 2, 3), (1, 2, 3, 0), (2, 3, 0, 1), (3, 0, 1, 2)]
    return Case(f"C4_{n}", [b, a], [ViewSpec.dense(0, (n,) * 4)] + [Av.permutedims(p) for p in perms], P_SUM4)


def case_c5(g=8, n=4096, dt=np.float64, seed=SEED):
    rng = np.random.default_rng(seed)
    a, out = randn(rng, g * n * n, dt), np.zeros(g, dt)
    dims = (g, n, n)
    return Case(f"C5_{g}x{n}", [out, a], [ViewSpec(0, 0, dims, (1, 0, 0)), ViewSpec.dense(1, dims)], P_ABS2, op=1,
                rtol=1e-6)
