"""KAT-6 physics on the GPU: damping / growth rates from full runs of the self-consistent Strang
stepper against the numerically solved kinetic dispersion relation; energy conservation of the scheme.
Tolerances are stated here: Landau gamma within 8 %, bump-on-tail gamma within 12 %, relative energy
drift < 1e-3 over 500 steps.  (The reference only plots these, scripts/bump_on_tail.jl:64-71.)"""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_landau_and_bump_on_tail_rates():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import physics_checks
    out = physics_checks.main(int(2e7))
    la, bs, bo = out["landau"], out["bump_on_tail_small_eps"], out["bump_on_tail"]
    assert abs(la["gamma_theory"] + 0.1533) < 1e-3                       # dispersion solver sanity
    assert abs(la["gamma_fit"] - la["gamma_theory"]) < 0.08 * abs(la["gamma_theory"])
    assert la["energy_drift_rel"] < 1e-3
    assert abs(bs["gamma_fit"] - bs["gamma_theory"]) < 0.12 * bs["gamma_theory"]
    assert bo["W_max"] > 10 * bo["W_first"] and bo["energy_drift_rel"] < 5e-3
