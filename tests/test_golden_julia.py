"""Fixtures written by the REAL Strided.jl (julia/make_fixtures.jl, run on a machine that has Julia) replayed through
the C restatement of the reference (CPU suite) and through the CUDA path (GPU suite).  This is the pin of the oracle to
reference-produced output; with no fixtures committed the tests skip and parity stays "unpinned".

A schema self-test runs in any case: a fixture in the Julia writer's format is synthesised from the NumPy oracle,
written to a temporary directory and consumed by the same loader, so the consumer cannot rot while no Julia fixture
exists."""
import glob
import os

import numpy as np
import pytest

from helpers import Case, ViewSpec, case_c2, case_c5

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden_julia", "*.npz")))


def load_julia_fixture(path):
    """-> (Case, expected output parent, exact).  Format: julia/make_fixtures.jl `save`."""
    z = np.load(path)
    parents = [np.ascontiguousarray(z[f"parent{i}"]).reshape(-1) for i in range(int(z["nparents"]))]
    dims = tuple(int(x) for x in np.atleast_1d(z["dims"]))
    vstr = np.asarray(z["vstrides"], dtype=np.int64).reshape(len(np.atleast_1d(z["vparent"])), -1)
    views = [ViewSpec(int(p), int(o), dims, tuple(int(s) for s in st), bool(cj))
             for p, o, st, cj in zip(np.atleast_1d(z["vparent"]), np.atleast_1d(z["voffset"]), vstr, np.atleast_1d(z["vconj"]))]
    toks = np.asarray(z["tokens"], dtype=np.float64).reshape(-1, 4)
    tokens = [(int(t[0]), int(t[1]), float(t[2]), float(t[3])) for t in toks]
    init = complex(float(z["init_re"]), float(z["init_im"]))
    rtol = float(z["rtol"])
    name = os.path.splitext(os.path.basename(path))[0]
    case = Case(name, parents, views, tokens, int(z["op"]), int(z["initop"]), init if init.imag else init.real, dims,
                None if rtol < 0 else rtol)
    expected = np.ascontiguousarray(z["expected"]).reshape(-1)
    assert expected.dtype == parents[views[0].parent].dtype and expected.size == parents[views[0].parent].size
    return case, expected, bool(int(z["exact"]))


def write_like_julia(path, case, expected, exact):
    """the same keys / shapes as julia/make_fixtures.jl `save` (NPZ.jl writes plain .npy members, column-major data is flat)"""
    d = {"nparents": np.int64(len(case.parents)), "dims": np.array(case.dims, dtype=np.int64),
         "vparent": np.array([v.parent for v in case.views], dtype=np.int64), "voffset": np.array([v.offset for v in case.views], dtype=np.int64),
         "vconj": np.array([int(v.conj) for v in case.views], dtype=np.int64),
         "vstrides": np.array([list(v.strides) for v in case.views], dtype=np.int64).reshape(len(case.views), -1),
         "tokens": np.array(case.tokens, dtype=np.float64).reshape(-1, 4), "op": np.int64(case.op), "initop": np.int64(case.initop),
         "init_re": np.float64(complex(case.init).real), "init_im": np.float64(complex(case.init).imag), "expected": expected,
         "exact": np.int64(exact), "rtol": np.float64(-1.0 if case.rtol is None else case.rtol)}
    for i, p in enumerate(case.parents):
        d[f"parent{i}"] = p
    np.savez(path, **d)


def test_fixture_schema_roundtrip(tmp_path):
    for i, (c, exact) in enumerate(((case_c2(48), True), (case_c5(4, 32), False))):
        path = os.path.join(tmp_path, f"synth_{i}.npz")
        write_like_julia(path, c, c.expected(), exact)
        c2, want, ex = load_julia_fixture(path)
        assert ex == exact and c2.dims == c.dims and c2.tokens == c.tokens
        c2.assert_close(c2.run_ref(2), want, exact=ex)


@pytest.mark.skipif(not FIXTURES, reason="no Julia-produced fixtures committed (tests/golden_julia/README.md): parity unpinned")
@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: os.path.basename(p))
def test_restated_reference_matches_julia(path):
    case, want, exact = load_julia_fixture(path)
    for nthreads in (1, 3):
        case.assert_close(case.run_ref(nthreads), want, exact=exact)


@pytest.mark.gpu
@pytest.mark.skipif(not FIXTURES, reason="no Julia-produced fixtures committed (tests/golden_julia/README.md): parity unpinned")
@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: os.path.basename(p))
def test_cuda_matches_julia(path):
    case, want, exact = load_julia_fixture(path)
    case.assert_close(case.run_gpu("device"), want, exact=exact)
