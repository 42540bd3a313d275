"""`losses.py` of this package against values produced by the REAL reference module
(`tests/golden/make_golden.py::golden_losses`, which imports `/root/reference/losses.py`)."""
import json
import os

import numpy as np
import pytest

import plfem_b200 as P
from plfem_b200.losses import EnhancedLossCalculator, LossCalculator
from golden.make_golden import synthetic_vectorial_modes

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "losses.json")))


def test_reference_selfcheck_numbers():
    """XT = -25.31 dB and PDL = 0.878 dB: what `python losses.py` prints in the reference (`losses.py:1252-1259`)."""
    m7 = synthetic_vectorial_modes(7, 42)
    assert abs(GOLD["selfcheck"]["xt"] + 25.31) < 5e-3 and abs(GOLD["selfcheck"]["pdl"] - 0.878) < 5e-4
    assert EnhancedLossCalculator._calculate_crosstalk(m7) == pytest.approx(GOLD["selfcheck"]["xt"], rel=1e-13)
    assert EnhancedLossCalculator._calculate_pdl_vectorial(m7) == pytest.approx(GOLD["selfcheck"]["pdl"], rel=1e-13)


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"{c['n_cores']}c-{c['direction']}-{int(c['wavelength_nm'])}")
def test_physical_losses_match_reference(case):
    g = P.MCFGeometry(case["n_cores"], 8.0, 1.5, 1.535, 1.0)
    modes = synthetic_vectorial_modes(case["n_modes"], case["seed"])
    res = LossCalculator.calculate_physical_losses(modes, g, case["direction"], case["wavelength_nm"])
    assert set(res) == set(case["result"])
    for k, v in case["result"].items():
        if isinstance(v, float):
            assert res[k] == pytest.approx(v, rel=1e-12, abs=1e-14), k
        else:
            assert res[k] == v, k
    dp = LossCalculator._build_design_params(modes, g, case["wavelength_nm"])
    for k, v in case["design"].items():
        mine = getattr(dp, k)
        if isinstance(v, float) and not isinstance(mine, (str, bool)):
            assert float(mine) == pytest.approx(v, rel=1e-12), k
        else:
            assert mine == v, k


def test_scalar_modes_are_refused_and_empty_lists_fail_cleanly():
    g = P.MCFGeometry(7, 8.0, 1.5, 1.535, 1.0)
    assert LossCalculator.calculate_physical_losses([], g)["success"] is False
    assert LossCalculator.calculate_physical_losses([{"n_eff": 1.2, "is_vectorial": False}], g)["success"] is False
    assert EnhancedLossCalculator._calculate_crosstalk([]) == -70.0
    assert EnhancedLossCalculator._calculate_crosstalk_vectorial([{"n_eff": 1.2}]) == -25.0
