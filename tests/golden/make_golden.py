"""Generates tests/golden/*.npz: small seeded inputs + expected outputs for the BASELINE recipes and a sample
of the reference's test recipes.  The expected values come from the NumPy semantic oracle (oracle/semantic.py)
-- NOT from Julia: the reference cannot run in this image (no julia) and stores no golden vectors itself, so
these fixtures pin the oracles and the CUDA path to each other and to Base-`Array` semantics over time.

    python tests/golden/make_golden.py        # rewrites the fixtures (deterministic: seed 1234)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402
from helpers import case_c1, case_c2, case_c3, case_c4, case_c5  # noqa: E402


def main():
    todo = [(case_c1(40), True), (case_c2(48), True), (case_c3(6), True), (case_c4(6), True), (case_c5(8, 24), False)]
    todo += [(c, True) for c in cases.inplace_matrix_cases(24) if c.name in ("adjoint!_complex128", "conj!_complex64")]
    pick = {"axpby_float64_N4", "lambda3_float32_N3", "bc3_complex64", "initop_scale_float64",
            "count_neg_perm_float64", "negstride_float64", "odd103_float32", "matmul_initop", "max_abs_perm_complex128"}
    for c in cases.all_cases(0.3):
        if c.name in pick:
            exact = c.op == 0 and c.parents[c.views[0].parent].dtype.kind != "c" and "lambda3" not in c.name
            todo.append((c, exact))
    # the alias-fused orbit path (three rotated views of one parent; 16^3 tiles): smallest shape that takes it
    from test_orbit import orbit_cases  # noqa: E402
    todo += [(c, True) for c, _ in orbit_cases() if c.name == "orbit_sum3_float32"]
    for c, exact in todo:
        z = c.to_npz(c.expected(), exact)
        np.savez_compressed(os.path.join(HERE, c.name.replace("!", "_") + ".npz"), **z)
        print("wrote", c.name, sum(v.nbytes for v in z.values() if hasattr(v, "nbytes")), "bytes")


if __name__ == "__main__":
    main()
