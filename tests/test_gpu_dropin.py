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


def test_fit_batch_equals_individual_fits():
    """`fit_batch` (additive helper over tame_fit_batch, BASELINE config 5): every object ends up exactly where its own
    `fit(max_iter, tolerance)` would have put it -- history (incl. the per-fit early stop of base.py:183-203) and state."""
    from src.models import TemporalAMEModel
    from src.inference import TemporalAMENaiveMFVI, TemporalAMEStructuredMFVI, fit_batch
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        def make():
            vis = []
            for (n, T, ar, rho) in [(10, 5, 0.8, 0.5), (24, 3, 0.5, 0.0), (70, 4, 0.9, 0.8), (33, 9, 0.6, 0.3)]:
                model = TemporalAMEModel(n_nodes=n, n_time=T, latent_dim=2, ar_coefficient=ar, rho_dyadic=rho, seed=42)
                model.generate_data()
                vis.append(TemporalAMENaiveMFVI(model, learning_rate=0.01, seed=42))
                vis.append(TemporalAMEStructuredMFVI(model, factorization="good", learning_rate=1e-5 if n == 10 else 0.01, seed=42))
            return vis
        a, b = make(), make()
        for vi in a:
            vi.fit(max_iter=12, tolerance=1.09e-3, verbose=False)
        hist = fit_batch(b, max_iter=12, tolerance=1.09e-3)
        assert len(hist) == len(b)
        stopped = 0
        for va, vb in zip(a, b):
            ea, eb = np.array(va.history["elbo"]), np.array(vb.history["elbo"])
            assert len(ea) == len(eb), "per-fit early stop differs"
            stopped += len(ea) < 12
            assert np.all(np.abs(ea - eb) <= 1e-9 * np.abs(ea))
            ma, mb = np.array(va.history["reconstruction_error"]), np.array(vb.history["reconstruction_error"])
            assert np.all(np.abs(ma - mb) <= 1e-9 * np.abs(ma))
            assert torch.allclose(va.X_mean, vb.X_mean, rtol=1e-9, atol=1e-12)
            assert torch.allclose(va.X_cov, vb.X_cov, rtol=1e-9, atol=1e-12)
        assert stopped >= 1          # the tiny-step fit stops early in both paths
        # a batch-fitted object keeps working like any other: continue it on its own
        b[1].fit(max_iter=2, tolerance=0.0, verbose=False)
        assert len(b[1].history["elbo"]) == len(a[1].history["elbo"]) + 2
    finally:
        torch.set_default_dtype(old)


def test_resume_after_pickle_is_exact_and_in_place_edits_are_seen(temporal_data):
    """The device state is FP64 while the handed-out tensors keep the reference's default dtype (float32): a pickled /
    resumed fit continues from the exact state (same trace as an uninterrupted fit), and an in-place edit of `vi.X_mean`
    -- the reference mutates its tensors in place -- reaches the device before the next sweep."""
    from src.inference import TemporalAMEStructuredMFVI
    model = temporal_data["model"]
    a = TemporalAMEStructuredMFVI(model, factorization="good", learning_rate=0.3, seed=42)
    ha = list(a.fit(max_iter=5, tolerance=0.0, verbose=False)["elbo"])
    b = TemporalAMEStructuredMFVI(model, factorization="good", learning_rate=0.3, seed=42)
    b.fit(max_iter=3, tolerance=0.0, verbose=False)
    assert b.X_mean.dtype == torch.float32
    c = pickle.loads(pickle.dumps(b))
    hc = list(c.fit(max_iter=2, tolerance=0.0, verbose=False)["elbo"])
    assert len(hc) == 5
    assert np.all(np.abs(np.array(hc) - np.array(ha)) <= 1e-12 * np.abs(np.array(ha))), (hc, ha)
    e0 = c._compute_elbo()
    c.X_mean.mul_(0.5)                       # in place, on the tensor the property handed out
    e1 = c._compute_elbo()
    assert abs(e1 - e0) > 1e-6 * abs(e0)
