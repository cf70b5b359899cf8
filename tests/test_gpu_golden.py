"""GPU: the CUDA path against the committed golden vectors (no oracle library involved at run time):
fields within 1e-12 relative per step, exact zero masks, identical solver iteration counts."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden as MG  # noqa: E402
from parity import load_pop, pop_global, relerr, c  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(MG.CASES))
def test_cuda_path_matches_golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    p = load_pop(MG.build_case(name))
    try:
        its = []
        for ts in MG.steps_of(name):
            p.step(ts)
            its.append(p.solvers_get_diagnostics()[0])
        assert its == g["solver_iterations"].tolist()
        for n in MG.FIELDS_OUT:
            a = pop_global(p, n, c.TIME_CUR)
            assert relerr(a, g[n]) <= 2.0e-12, (n, relerr(a, g[n]))
            assert np.array_equal(a == 0.0, g[n] == 0.0), n
    finally:
        p.finalize()
