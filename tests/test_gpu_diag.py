"""Parity of the device diagnostics (tame_contributions / tame_uv_correlation / tame_state_mse through the
reference-named functions of src.utils.diagnostics) with the golden outputs of the unmodified reference
(tests/golden/diag.npz) and with the oracle at benchmark sizes.  Tolerance rel 1e-9 (correlations: abs 1e-9)."""
import os

import numpy as np
import pytest
import torch

from oracle import diag_oracle as do

pytestmark = pytest.mark.gpu
TOL = 1e-9
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "diag.npz"))
CASES = [str(c) for c in G["cases"]]


@pytest.fixture(autouse=True)
def _float64_default():
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


@pytest.mark.parametrize("case", CASES)
def test_diagnostics_match_reference_golden(case):
    from src.utils import (compute_additive_contribution, compute_contribution_ratio, compute_multiplicative_contribution,
                           compute_state_prediction_error, compute_temporal_contributions, compute_uv_correlation_over_time,
                           compute_uv_product_correlation)
    r = int(G[f"{case}_r"])
    Xe, Xt = _t(G[f"{case}_est"]), _t(G[f"{case}_true"])
    for tag, excl in (("excl", True), ("incl", False)):
        add, mul = compute_temporal_contributions(Xe, r, exclude_diagonal=excl)
        assert add.shape == (Xe.shape[1],) and not add.is_cuda
        assert np.allclose(add.numpy(), G[f"{case}_add_{tag}"], rtol=TOL, atol=0)
        assert np.allclose(mul.numpy(), G[f"{case}_mul_{tag}"], rtol=TOL, atol=0)
        a0 = compute_additive_contribution(Xe[:, 0, :2], excl)
        m0 = compute_multiplicative_contribution(Xe[:, 0, 2:].cuda(), excl)
        assert abs(a0 - G[f"{case}_add_{tag}"][0]) <= TOL * a0 and abs(m0 - G[f"{case}_mul_{tag}"][0]) <= TOL * m0
    ratio = compute_contribution_ratio(Xe[:, 0, :2], Xe[:, 0, 2:])
    assert abs(ratio - float(G[f"{case}_ratio0"])) <= TOL * ratio
    mse = compute_state_prediction_error(Xt, Xe)
    assert abs(mse - float(G[f"{case}_state_mse"])) <= TOL * mse
    assert np.allclose(compute_uv_correlation_over_time(Xe, Xt, r).numpy(), G[f"{case}_uvcorr_t"], rtol=0, atol=TOL)
    assert abs(compute_uv_product_correlation(Xe[:, 0, 2:], Xt[:, 0, 2:], r) - float(G[f"{case}_uvcorr0"])) <= TOL


def test_edge_cases():
    from src.utils.diagnostics import compute_contribution_ratio, compute_temporal_contributions
    A = torch.randn(9, 2)
    assert compute_contribution_ratio(A, torch.zeros(9, 4)) == float("inf")       # diagnostics.py:247-248
    with pytest.raises(ValueError):
        compute_temporal_contributions(torch.zeros(4, 3, 6), latent_dim=3)


@pytest.mark.parametrize("shape", [(1024, 64, 4), (777, 19, 7), (4096, 128, 8)])
def test_against_oracle_at_benchmark_sizes(shape):
    """Literal n x n oracle vs the Gram-identity kernels at config 3 and at half of config 4's n (the oracle's n^2 T
    products stay within seconds)."""
    from tame_b200 import diagnostics as dg
    n, T, r = shape
    d = 2 + 2 * r
    g = torch.Generator(device="cuda").manual_seed(n + T)
    Xt = torch.randn(n, T, d, generator=g, dtype=torch.float64, device="cuda")
    Xe = 0.8 * Xt + 0.4 * torch.randn(n, T, d, generator=g, dtype=torch.float64, device="cuda") + 0.05
    add, mul = dg._contrib_device(Xe, r, True)
    corr = dg.compute_uv_correlation_over_time(Xe, Xt, r).numpy()
    ts = sorted({0, 1, T // 2, T - 1})
    xe, xt = Xe[:, ts].cpu().numpy(), Xt[:, ts].cpu().numpy()
    oadd, omul = do.temporal_contributions(xe, r, True)
    ocorr = do.uv_correlation_over_time(xe, xt, r)
    assert np.allclose(add.cpu().numpy()[ts], oadd, rtol=TOL, atol=0)
    assert np.allclose(mul.cpu().numpy()[ts], omul, rtol=TOL, atol=0)
    assert np.allclose(corr[ts], ocorr, rtol=0, atol=TOL)
