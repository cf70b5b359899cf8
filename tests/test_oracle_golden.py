"""CPU: the oracle reproduces the committed golden vectors bit for bit (tools/make_golden.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden as MG  # noqa: E402
from parity import load_oracle, oracle_global, c  # noqa: E402


@pytest.mark.parametrize("name", list(MG.CASES))
def test_oracle_matches_golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    o = load_oracle(MG.build_case(name))
    its = []
    for ts in MG.steps_of(name):
        assert o.step(ts) == 0
        its.append(o.solver_diag()[0])
    assert its == g["solver_iterations"].tolist()
    for n in MG.FIELDS_OUT:
        assert np.array_equal(oracle_global(o, n, c.TIME_CUR), g[n]), n
