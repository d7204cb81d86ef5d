"""CPU suite: the CUDA kernel BODIES (csrc/map_tile.hpp, csrc/reduce_tile.hpp) and the C++ planner, executed
for every (block, thread) on the CPU by tests/emul/.  This checks the index machinery -- per-operand load
orders, offset functionals, staging slots, edge masks, split reductions -- without a GPU.  It is not a
product path: the emulator library is built only here."""
import numpy as np
import pytest

import cases
from helpers import case_c1, case_c2, case_c3, case_c4, case_c5


@pytest.mark.parametrize("case", cases.all_cases(0.3), ids=lambda c: c.name)
def test_emulated_kernel_matches_semantic_oracle(case):
    want = case.expected()
    case.assert_close(case.run_emul(), want)
    case.assert_close(case.run_emul(grid_limit=2), want)  # few persistent CTAs: exercises the tile loop


def test_emulated_baseline_configs():
    for c in (case_c1(150), case_c2(260), case_c3(16), case_c4(16)):
        c.assert_close(c.run_emul(), exact=True)
    c5 = case_c5(8, 96)
    c5.assert_close(c5.run_emul())


def test_golden_fixtures_through_the_emulated_kernels():
    """the committed fixtures (tests/golden/) through the planner + CPU thread-grid emulation of the kernel bodies"""
    import os
    from helpers import Case, ROOT
    gdir = os.path.join(ROOT, "tests", "golden")
    for f in sorted(os.listdir(gdir)):
        if not f.endswith(".npz"):
            continue
        z = np.load(os.path.join(gdir, f), allow_pickle=False)
        case = Case.from_npz(z)
        case.assert_close(case.run_emul(), z["expected"], exact=bool(z["exact"]))


def test_emulated_lsu_tile_records(monkeypatch):
    """per-tile records of the LSU map kernel (MapParams::lsu_desc): the table is built from 512 tiles on; here it is forced
    on every multi-dim map plan and must give the results of the in-kernel decode (edge tiles, shifted tiles, in-place
    updates that drop the table at bind time, tile orders)."""
    monkeypatch.setenv("SB_LSU_DESC_MIN", "1")
    monkeypatch.setenv("SB_NO_TMA", "1")
    seen = 0
    for case in cases.all_cases(0.3):
        if case.op != 0:
            continue
        if case.plan().get("lsu_desc"):
            seen += 1
        case.assert_close(case.run_emul(grid_limit=3), case.expected())
    assert seen > 50
