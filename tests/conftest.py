import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def pytest_sessionstart(session):
    """Safety net: the CPU suite needs the planner inside libstrided_b200.so (sb_plan_describe) and the oracle library;
    build them if a fresh checkout has not run __graft_entry__.build() yet (nvcc cross-compiles without a GPU)."""
    lib = os.path.join(ROOT, "strided.jl_b200", "libstrided_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__ as g
        g.build()
