"""The oracle (oracle/tame_oracle.py) against the golden vectors produced by the reference itself
(tests/golden/make_golden.py, float64).  CPU only."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, GOLDEN_LONG, METHODS, golden_constants, load_golden, rel_err
from oracle import tame_oracle as orc

# rel 1e-9 is north_star's tolerance; the oracle is expected to sit far inside it.
TOL = 1e-9
# 'bad' with lr=1 is a divergent iteration (ELBO -> -4e4 in the reference itself): rounding differences are
# amplified by orders of magnitude per sweep, so only its first iterations are comparable.
UNSTABLE = {("conftest_lr1", "bad"), ("r3_rho08", "bad"), ("r1_T1", "bad"), ("r4_T2", "bad")}


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_model_constants_match_reference(case):
    g = load_golden(case)
    kw = eval(str(g["model_kwargs"]))  # fixture written by make_golden.py
    c = orc.model_constants(kw["n_nodes"], kw["n_time"], kw.get("latent_dim", 2), kw.get("ar_coefficient", 0.8),
                            kw.get("rho_additive", 0.5), kw.get("rho_multiplicative", 0.3), kw.get("rho_dyadic", 0.5))
    for k in ("R", "Sigma", "Psi", "Phi", "Q"):
        assert rel_err(c[k], g[k]) < 1e-15, k
    assert rel_err(c["R_inv"], g["R_inv"]) < 1e-13


@pytest.mark.parametrize("meth", METHODS)
@pytest.mark.parametrize("case", [c for c in GOLDEN_CASES if c not in ("config2",)])
def test_fit_trace_matches_reference(case, meth):
    g = load_golden(case)
    c = golden_constants(g)
    Xm, Xc = g[f"{meth}_init_mean"].copy(), g[f"{meth}_init_cov"].copy()
    el, ms = orc.fit(g["Y"], Xm, Xc, c, float(g["lr"]), orc.MODE_OF[meth], int(g["max_iter"]), float(g["tolerance"]))
    ref_el, ref_ms = g[f"{meth}_elbo"], g[f"{meth}_mse"]
    assert len(el) == len(ref_el)
    if (case, meth) in UNSTABLE:
        assert abs(el[0] - ref_el[0]) <= TOL * abs(ref_el[0])
        assert abs(ms[0] - ref_ms[0]) <= TOL * abs(ref_ms[0])
        return
    assert np.all(np.abs(el - ref_el) <= TOL * np.abs(ref_el))
    assert np.all(np.abs(ms - ref_ms) <= TOL * np.abs(ref_ms))
    assert rel_err(Xm, g[f"{meth}_final_mean"]) < TOL
    assert rel_err(Xc, g[f"{meth}_final_cov"]) < TOL
    parts = orc.elbo_parts(g["Y"], Xm, Xc, c, orc.MODE_OF[meth])
    assert np.all(np.abs(np.array(parts) - g[f"{meth}_parts"]) <= TOL * np.abs(g[f"{meth}_parts"]))


@pytest.mark.parametrize("meth", METHODS)
def test_config2_first_iterations(meth):
    g = load_golden("config2")
    c = golden_constants(g)
    Xm, Xc = g[f"{meth}_init_mean"].copy(), g[f"{meth}_init_cov"].copy()
    el, ms = orc.fit(g["Y"], Xm, Xc, c, float(g["lr"]), orc.MODE_OF[meth], int(g["max_iter"]), 0.0)
    assert np.all(np.abs(el - g[f"{meth}_elbo"]) <= TOL * np.abs(g[f"{meth}_elbo"]))
    assert np.all(np.abs(ms - g[f"{meth}_mse"]) <= TOL * np.abs(g[f"{meth}_mse"]))
    assert rel_err(Xm, g[f"{meth}_final_mean"]) < TOL
    assert rel_err(Xc, g[f"{meth}_final_cov"]) < TOL


@pytest.mark.parametrize("meth", METHODS)
@pytest.mark.parametrize("case", GOLDEN_LONG)
def test_config2_fifty_iterations(case, meth):
    """BASELINE config 2 (three_way_conparison: n=50, T=20, r=2, lr 0.01), 50 iterations of every method: the whole ELBO
    and MSE traces and the final state of the reference, against the oracle's BLAS-vectorised literal sweep."""
    g = load_golden(case)
    c = golden_constants(g)
    mode = orc.MODE_OF[meth]
    Xm, Xc = g[f"{meth}_init_mean"].copy(), g[f"{meth}_init_cov"].copy()
    el, ms = [], []
    for _ in range(int(g["max_iter"])):
        orc.sweep_fast(g["Y"], Xm, Xc, c, float(g["lr"]), mode)
        el.append(orc.elbo(g["Y"], Xm, Xc, c, mode))
        ms.append(orc.reconstruction_mse(g["Y"], Xm, c))
    el, ms = np.array(el), np.array(ms)
    assert np.all(np.abs(el - g[f"{meth}_elbo"]) <= TOL * np.abs(g[f"{meth}_elbo"]))
    assert np.all(np.abs(ms - g[f"{meth}_mse"]) <= TOL * np.abs(g[f"{meth}_mse"]))
    assert rel_err(Xm, g[f"{meth}_final_mean"]) < TOL
    assert rel_err(Xc, g[f"{meth}_final_cov"]) < TOL


@pytest.mark.parametrize("meth", METHODS)
@pytest.mark.parametrize("block", [1, 3, 4, 64])
def test_blocked_and_fast_sweeps_keep_the_schedule(meth, block):
    """The restructured sweep (totals + static upper part + pushes + inline window) and the BLAS
    variant reproduce the literal Gauss-Seidel order."""
    g = load_golden("r3_rho08")
    c = golden_constants(g)
    mode = orc.MODE_OF[meth]
    lr = 0.3
    A = (g[f"{meth}_init_mean"].copy(), g[f"{meth}_init_cov"].copy())
    B = (A[0].copy(), A[1].copy())
    F = (A[0].copy(), A[1].copy())
    for _ in range(3):
        orc.sweep(g["Y"], A[0], A[1], c, lr, mode)
        orc.sweep_blocked(g["Y"], B[0], B[1], c, lr, mode, block=block)
        orc.sweep_fast(g["Y"], F[0], F[1], c, lr, mode)
    assert rel_err(B[0], A[0]) < 1e-11 and rel_err(B[1], A[1]) < 1e-11
    assert rel_err(F[0], A[0]) < 1e-11 and rel_err(F[1], A[1]) < 1e-11


def test_golden_data_properties():
    """Reference data invariants the kernels rely on or must not rely on (test_models.py:138-142)."""
    g = load_golden("config1")
    Y = g["Y"]
    n = Y.shape[0]
    assert np.all(Y[np.arange(n), np.arange(n)] == 0.0)
    assert np.array_equal(Y[:, :, :, 1], np.swapaxes(Y[:, :, :, 0], 0, 1))
    assert abs(Y.sum() - 1272.748187091072) < 1e-9
    assert abs(g["good_init_mean"].sum() - 2.303365457538453) < 1e-12
    assert abs(g["good_init_cov"].sum() - 541.2638460976557) < 1e-9
