"""`fit_batch(..., devices=[0, 1])`: independent fits dealt over two GPUs (needs >= 2 visible GPUs; skipped otherwise).
Kept in a file that sorts last so that the single-GPU parity suites always run first."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_fit_batch_over_two_devices_equals_one_device():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "python-temporal-ame-svi_b200"))
    from src.models import TemporalAMEModel
    from src.inference import TemporalAMENaiveMFVI, TemporalAMEStructuredMFVI, fit_batch
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        def make():
            vis = []
            for (n, T, ar, rho) in [(10, 5, 0.8, 0.5), (24, 3, 0.5, 0.0), (70, 4, 0.9, 0.8), (33, 9, 0.6, 0.3), (48, 6, 0.7, 0.2)]:
                model = TemporalAMEModel(n_nodes=n, n_time=T, latent_dim=2, ar_coefficient=ar, rho_dyadic=rho, seed=42)
                model.generate_data()
                vis.append(TemporalAMENaiveMFVI(model, learning_rate=0.01, seed=42))
                vis.append(TemporalAMEStructuredMFVI(model, factorization="good", learning_rate=0.01, seed=42))
            return vis
        a, b = make(), make()
        fit_batch(a, max_iter=8, tolerance=0.0, device="cuda:0")
        hist = fit_batch(b, max_iter=8, tolerance=0.0, devices=[0, 1])
        assert len(hist) == len(b)
        for va, vb in zip(a, b):
            ea, eb = np.array(va.history["elbo"]), np.array(vb.history["elbo"])
            assert len(ea) == len(eb) == 8
            assert np.all(np.abs(ea - eb) <= 1e-9 * np.abs(ea))
            assert torch.allclose(va.X_mean, vb.X_mean, rtol=1e-9, atol=1e-12)
            assert torch.allclose(va.X_cov, vb.X_cov, rtol=1e-9, atol=1e-12)
    finally:
        torch.set_default_dtype(old)
