"""The diagnostics oracle (oracle/diag_oracle.py) against outputs of the unmodified reference (tests/golden/diag.npz)."""
import os

import numpy as np
import pytest

from oracle import diag_oracle as do

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "diag.npz"))
CASES = [str(c) for c in G["cases"]]
TOL = 1e-12


@pytest.mark.parametrize("case", CASES)
def test_diagnostics_oracle_matches_reference(case):
    r = int(G[f"{case}_r"])
    Xe, Xt = G[f"{case}_est"], G[f"{case}_true"]
    for tag, excl in (("excl", True), ("incl", False)):
        add, mul = do.temporal_contributions(Xe, r, excl)
        assert np.allclose(add, G[f"{case}_add_{tag}"], rtol=TOL, atol=0) and np.allclose(mul, G[f"{case}_mul_{tag}"], rtol=TOL, atol=0)
    assert abs(do.contribution_ratio(Xe[:, 0, :2], Xe[:, 0, 2:]) - float(G[f"{case}_ratio0"])) <= TOL * float(G[f"{case}_ratio0"])
    assert abs(do.state_prediction_error(Xt, Xe) - float(G[f"{case}_state_mse"])) <= TOL * float(G[f"{case}_state_mse"])
    assert np.allclose(do.uv_correlation_over_time(Xe, Xt, r), G[f"{case}_uvcorr_t"], rtol=0, atol=1e-12)
    assert abs(do.uv_product_correlation(Xe[:, 0, 2:], Xt[:, 0, 2:], r) - float(G[f"{case}_uvcorr0"])) <= 1e-12
