"""CPU: pin the oracle (oracle/sph_oracle.c) against fixtures generated from the real
reference build (tests/golden/make_golden.py), including the six known-answer decks of
examples/USER/sph/multiphase_two_atoms whose closed-form values are in SURVEY.md 4."""
import numpy as np
import pytest

import cases
import harness

TOL_STEP = 1e-12     # one force evaluation: oracle restates the reference operation by operation
ALL = list(cases.CASES)


@pytest.mark.parametrize("name", ALL)
def test_oracle_matches_reference(name):
    e0, eN = harness.run_case(harness.oracle_sim, name, tol_step=TOL_STEP, tol_traj=1e-10)


def test_known_answers():
    """closed-form values of the Maxima scripts, as printed by the reference"""
    g = harness.load_golden("kat_rhosum_multiphase")
    assert np.allclose(g["s0_rho"], [17.320034294651357243, 5.6179934561675270999, 8.490241626116999285], rtol=1e-15)
    g = harness.load_golden("kat_taitwater_multiphase")
    assert np.allclose(g["s0_f"][:, 0], [-12.589404678167504414, 16.123080287781466069, -3.5336756096139603223], rtol=1e-14)
    g = harness.load_golden("kat_colorgradient")
    cg = np.linalg.norm(g["s0_colorgradient"], axis=1)
    assert np.allclose(cg[:2], [6.1824490426522382691, 20.143047485068006353], rtol=1e-14)
    g = harness.load_golden("kat_surfacetension")
    assert np.allclose(g["s0_f"][0], [-27.049911225862341979, 20.287336097181302819, -5.1905181562649258887e-05], rtol=1e-10)
    g = harness.load_golden("kat_heatconduction_phasechange")
    assert np.allclose(g["s0_de"], [24.637185190625263687, -12.318592595312631843], rtol=1e-14)
    g = harness.load_golden("kat_phase_change")
    assert len(g["sN_type"]) == 3 and np.allclose(g["sN_rmass"], [9, 2, 1]) and np.allclose(g["sN_e"], [11.1111111111111, 0.5, 0.5])
