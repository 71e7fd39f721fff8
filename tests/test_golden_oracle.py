"""Pins the oracle against golden vectors produced by the reference's own source (CPU only)."""
import numpy as np
import pytest

import golden_cases as gc
from conftest import rel_l2


@pytest.mark.parametrize("case", gc.CASES, ids=[c["name"] for c in gc.CASES])
def test_oracle_reproduces_reference_outputs(case):
    x, y, p = gc.arrays(case["name"])
    out = gc.oracle_eval(case, x, p)
    assert out.shape == y.shape
    assert rel_l2(out, y) < 1e-11


def test_parameter_counts_from_the_reference_itself():
    import pdequinox_b200 as pdeqx
    for name, n in gc.PARAMETER_COUNTS.items():
        if name == "ClassicFNO_2d_h64_m16":
            m = pdeqx.ClassicFNO(2, 1, 1, hidden_channels=64, num_modes=16, key=0)
        else:
            cls, d = name.rsplit("_", 1)
            m = getattr(pdeqx, cls)(int(d[0]), 1, 1, key=0)
        assert pdeqx.count_parameters(m) == n, name
