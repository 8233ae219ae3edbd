"""Drop-in acceptance on the GPU: the behaviours the reference's own tests/test_inference.py pins (shapes, structure
kept during optimisation, history bookkeeping, convergence stop, verbose text, learning-rate sensitivity, method
differences), exercised through the reference's import paths (`src.models`, `src.inference`) in the reference's default
dtype (float32 tensors in, FP64 arithmetic inside)."""
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def temporal_data():
    # tests/conftest.py:35-43,70-78 of the reference
    from src.models import TemporalAMEModel
    torch.manual_seed(42)
    np.random.seed(42)
    model = TemporalAMEModel(n_nodes=10, n_time=5, latent_dim=2, ar_coefficient=0.8, seed=42)
    Y, X = model.generate_data(return_latents=True)
    return {"Y": Y, "X": X, "model": model}


def test_naive_fit_history_and_getters(temporal_data):
    from src.inference import TemporalAMENaiveMFVI
    model = temporal_data["model"]
    vi = TemporalAMENaiveMFVI(model, learning_rate=0.01)
    history = vi.fit(max_iter=5, verbose=False)
    assert set(history) >= {"elbo", "reconstruction_error"}
    assert len(history["elbo"]) == 5 and len(history["reconstruction_error"]) == 5
    assert all(np.isfinite(history["elbo"])) and history["elbo"][-1] > -1e10
    assert vi.get_variational_means().shape == (model.n, model.T, model.d)
    assert vi.get_variational_covariances().shape == (model.n, model.T, model.d, model.d)
    assert vi.get_elbo_history() is history["elbo"] and len(vi.get_reconstruction_history()) == 5
    assert vi.predict_forward(n_steps=3).shape == (model.n, 3, model.d)
    cov = vi.X_cov
    off = cov - torch.diag_embed(torch.diagonal(cov, dim1=-2, dim2=-1))
    assert torch.count_nonzero(off) == 0                        # naive covariances stay diagonal
    assert vi.X_mean.dtype == torch.float32                     # handed back in the dtype the state was created in


def test_structured_good_and_bad_keep_their_structure(temporal_data):
    from src.inference import TemporalAMEStructuredMFVI
    model = temporal_data["model"]
    good = TemporalAMEStructuredMFVI(model, factorization="good")
    assert good.get_factorization_type() == "good"
    hg = good.fit(max_iter=5, verbose=False)
    assert len(hg["elbo"]) == 5
    off = good.X_cov - torch.diag_embed(torch.diagonal(good.X_cov, dim1=-2, dim2=-1))
    assert torch.count_nonzero(off) > 0
    bad = TemporalAMEStructuredMFVI(model, factorization="bad")
    bad.fit(max_iter=5, verbose=False)
    assert torch.allclose(bad.X_cov[:, :, :2, 2:], torch.zeros(model.n, model.T, 2, model.d - 2), atol=1e-5)
    assert torch.allclose(bad.X_cov[:, :, 2:, :2], torch.zeros(model.n, model.T, model.d - 2, 2), atol=1e-5)
    with pytest.raises(ValueError):
        TemporalAMEStructuredMFVI(model, factorization="invalid")


def test_reproducibility_and_method_differences(temporal_data):
    from src.inference import TemporalAMENaiveMFVI, TemporalAMEStructuredMFVI
    model = temporal_data["model"]
    a = TemporalAMENaiveMFVI(model, seed=42)
    a.fit(max_iter=5, verbose=False)
    b = TemporalAMENaiveMFVI(model, seed=42)
    b.fit(max_iter=5, verbose=False)
    assert torch.equal(a.X_mean, b.X_mean)                      # deterministic kernels: bit-identical reruns
    g = TemporalAMEStructuredMFVI(model, factorization="good", seed=42)
    hg = g.fit(max_iter=10, verbose=False)
    bd = TemporalAMEStructuredMFVI(model, factorization="bad", seed=42)
    hb = bd.fit(max_iter=10, verbose=False)
    assert not np.isclose(hg["reconstruction_error"][-1], hb["reconstruction_error"][-1], atol=0.001)


def test_learning_rates_convergence_and_verbose(temporal_data, capsys):
    from src.inference import TemporalAMENaiveMFVI
    model = temporal_data["model"]
    small = TemporalAMENaiveMFVI(model, learning_rate=0.001).fit(max_iter=5, verbose=False)
    large = TemporalAMENaiveMFVI(model, learning_rate=0.1).fit(max_iter=5, verbose=False)
    assert not np.allclose(small["elbo"], large["elbo"], atol=1.0)
    vi = TemporalAMENaiveMFVI(model, learning_rate=0.001)
    hist = vi.fit(max_iter=100, tolerance=1e-3, verbose=False)
    assert 4 <= len(hist["elbo"]) <= 100                        # early stop needs >= 3 consecutive small changes
    vi2 = TemporalAMENaiveMFVI(model)
    vi2.fit(max_iter=5, verbose=True, check_every=1)
    out = capsys.readouterr().out
    assert "Starting TemporalAMENaiveMFVI optimization..." in out
    assert "Iter" in out and "ELBO" in out and "MSE" in out and "Reached maximum iterations" in out


def test_state_assignment_and_pickle_round_trip(temporal_data):
    """experiments/utils.py:99-102 pickles the vi object; user code may also assign X_mean / X_cov."""
    from src.inference import TemporalAMEStructuredMFVI
    model = temporal_data["model"]
    vi = TemporalAMEStructuredMFVI(model, factorization="good", learning_rate=0.5)
    vi.fit(max_iter=3, verbose=False)
    clone = pickle.loads(pickle.dumps(vi))
    assert torch.equal(clone.X_mean, vi.X_mean) and clone.history == vi.history
    h1 = vi.fit(max_iter=2, verbose=False)["elbo"][-2:]
    h2 = clone.fit(max_iter=2, verbose=False)["elbo"][-2:]      # the unpickled object continues on a fresh engine
    assert np.allclose(h1, h2, rtol=1e-6)                        # (state went through float32 host tensors)
    vi.X_mean = torch.zeros_like(vi.X_mean)
    e0 = vi._compute_elbo()
    vi._update_step()
    assert vi._compute_elbo() != e0 and torch.count_nonzero(vi.X_mean) > 0
