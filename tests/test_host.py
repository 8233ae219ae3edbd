"""Host-side logic (no GPU): the model mirror reproduces the reference's data and hyper-parameters, the VI
classes reproduce the reference's RNG-order-exact initialisation, the C-ABI library loads and exports every
symbol of include/tame_b200.h."""
import ctypes
import os
import pickle
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, METHODS, ROOT, load_golden


@pytest.fixture(autouse=True)
def _float64_default():
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def _model(case):
    from tame_b200 import TemporalAMEModel
    g = load_golden(case)
    kw = eval(str(g["model_kwargs"]))
    m = TemporalAMEModel(**kw)
    return g, m


def _vi(meth, model, lr):
    from tame_b200 import TemporalAMENaiveMFVI, TemporalAMEStructuredMFVI
    if meth == "naive":
        return TemporalAMENaiveMFVI(model, learning_rate=lr, seed=42)
    return TemporalAMEStructuredMFVI(model, factorization=meth, learning_rate=lr, seed=42)


@pytest.mark.parametrize("case", [c for c in GOLDEN_CASES if c != "config2"])
def test_model_mirror_reproduces_reference_data(case):
    g, m = _model(case)
    for k in ("R", "R_inv", "Sigma", "Psi", "Phi", "Q"):
        assert np.array_equal(getattr(m, k).numpy(), g[k]), k
    Y, X = m.generate_data(return_latents=True)
    assert Y.dtype == torch.float64
    assert np.array_equal(X.numpy(), g["X_true"])
    assert np.array_equal(Y.numpy(), g["Y"])


@pytest.mark.parametrize("meth", METHODS)
@pytest.mark.parametrize("case", ["conftest_lr1", "r3_rho08", "config1"])
def test_vi_init_is_rng_exact(case, meth):
    g, m = _model(case)
    m.generate_data()
    vi = _vi(meth, m, float(g["lr"]))
    assert np.array_equal(vi.X_mean.numpy(), g[f"{meth}_init_mean"])
    assert np.array_equal(vi.X_cov.numpy(), g[f"{meth}_init_cov"])
    assert vi.n == m.n and vi.T == m.T and vi.d == m.d and vi.r == m.r


def test_seed_argument_is_ignored_like_the_reference():
    from tame_b200 import TemporalAMEModel
    a = TemporalAMEModel(n_nodes=6, n_time=3, seed=1).generate_data()
    b = TemporalAMEModel(n_nodes=6, n_time=3, seed=2).generate_data()
    assert torch.equal(a, b)


def test_reference_style_imports_and_structure():
    """tests/test_inference.py:30-43,119-154 of the reference, minus the parts that need the GPU."""
    from src.models import TemporalAMEModel
    from src.inference import TemporalAMENaiveMFVI, TemporalAMEStructuredMFVI
    torch.set_default_dtype(torch.float32)          # the reference's default run
    model = TemporalAMEModel(n_nodes=10, n_time=5, latent_dim=2, ar_coefficient=0.8, seed=42)
    model.generate_data()
    vi = TemporalAMENaiveMFVI(model, learning_rate=0.01)
    assert vi.lr == 0.01 and vi.X_mean.shape == (10, 5, 6) and vi.X_cov.shape == (10, 5, 6, 6)
    off = vi.X_cov - torch.diag_embed(torch.diagonal(vi.X_cov, dim1=-2, dim2=-1))
    assert torch.allclose(off, torch.zeros_like(off), atol=1e-6)
    good = TemporalAMEStructuredMFVI(model, factorization="good")
    off = good.X_cov - torch.diag_embed(torch.diagonal(good.X_cov, dim1=-2, dim2=-1))
    assert all(not torch.allclose(off[i, t], torch.zeros(6, 6)) for i in range(10) for t in range(5))
    bad = TemporalAMEStructuredMFVI(model, factorization="bad")
    assert bad.get_factorization_type() == "bad"
    assert torch.count_nonzero(bad.X_cov[:, :, :2, 2:]) == 0 and torch.count_nonzero(bad.X_cov[:, :, 2:, :2]) == 0
    with pytest.raises(ValueError):
        TemporalAMEStructuredMFVI(model, factorization="invalid")


def test_vi_object_pickles_without_device_handles():
    from tame_b200 import TemporalAMEModel, TemporalAMEStructuredMFVI
    model = TemporalAMEModel(n_nodes=6, n_time=3)
    model.generate_data()
    vi = TemporalAMEStructuredMFVI(model, factorization="good", learning_rate=0.5)
    vi2 = pickle.loads(pickle.dumps(vi))
    assert torch.equal(vi2.X_mean, vi.X_mean) and torch.equal(vi2.X_cov, vi.X_cov) and vi2.lr == 0.5


def test_no_cpu_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tame_b200 import TemporalAMEModel, TemporalAMENaiveMFVI
    model = TemporalAMEModel(n_nodes=6, n_time=3)
    model.generate_data()
    vi = TemporalAMENaiveMFVI(model)
    with pytest.raises(RuntimeError):
        vi.fit(max_iter=1, verbose=False)


def test_post_fit_utilities_have_no_cpu_fallback_and_keep_reference_names():
    """src.utils mirrors the reference's alignment / diagnostics names (src/utils/__init__.py:30-50); without a GPU
    they raise instead of computing on the host."""
    import src.utils as U
    for name in ("procrustes_alignment", "align_signs", "align_latent_positions", "align_temporal_states",
                 "compute_alignment_error", "compute_correlation_after_alignment", "compute_additive_contribution",
                 "compute_multiplicative_contribution", "compute_temporal_contributions", "compute_contribution_ratio",
                 "compute_state_prediction_error", "compute_uv_product_correlation"):
        assert callable(getattr(U, name)), name
    x = torch.zeros(4, 3, 6)
    with pytest.raises(ValueError):                      # alignment.py:354-357: checked before any device work
        U.compute_alignment_error(x, x)
    err, same = U.compute_alignment_error(x, x + 1.0, latent_dim=2, align=False)      # no alignment -> plain MSE
    assert same is x and err == 1.0
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            U.align_temporal_states(x, x, latent_dim=2)
        with pytest.raises(RuntimeError):
            U.compute_temporal_contributions(x, 2)


def test_library_exports_every_declared_symbol():
    from tame_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tame_b200.h")).read()
    declared = set(re.findall(r"\b(tame_[A-Za-z_0-9]+)\s*\(", header))
    declared -= {"tame_handle", "tame_config"}
    assert declared, "no declarations found"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/tame_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert b"sm_100a" in lib.tame_version()
    assert ctypes.sizeof(_lib.TameConfig) == 120
