"""The alignment oracle (oracle/align_oracle.py) against outputs of the unmodified reference module
(tests/golden/align.npz, made by tests/golden/make_golden_align.py).  Tolerance 1e-12: same LAPACK SVD underneath."""
import os

import numpy as np
import pytest

from oracle import align_oracle as ao

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "align.npz"))
TEMPORAL = [str(c) for c in G["temporal_cases"]]
TOL = 1e-12


def _close(a, b, tol=TOL):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) <= tol * max(1.0, np.max(np.abs(b)))


@pytest.mark.parametrize("case", TEMPORAL)
def test_align_temporal_states(case):
    r, each = int(G[f"{case}_r"]), bool(G[f"{case}_each"])
    out = ao.align_temporal_states(G[f"{case}_est"], G[f"{case}_true"], r, align_each_time=each)
    assert _close(out, G[f"{case}_aligned"])
    if each:
        err, al = ao.compute_alignment_error(G[f"{case}_est"], G[f"{case}_true"], latent_dim=r)
        assert abs(err - float(G[f"{case}_error"])) <= TOL * abs(err) and _close(al, G[f"{case}_aligned"])
        err0, al0 = ao.compute_alignment_error(G[f"{case}_est"], G[f"{case}_true"], latent_dim=r, align=False)
        assert abs(err0 - float(G[f"{case}_error_noalign"])) <= TOL * abs(err0) and al0 is G[f"{case}_est"] or np.array_equal(al0, G[f"{case}_est"])
        corr = ao.compute_correlation_after_alignment(G[f"{case}_est"], G[f"{case}_true"], latent_dim=r)
        assert abs(corr - float(G[f"{case}_corr"])) <= 1e-12


@pytest.mark.parametrize("case", ["s_r2", "s_r8"])
def test_static_states(case):
    r = int(G[f"{case}_r"])
    e, t = G[f"{case}_est"], G[f"{case}_true"]
    err, al = ao.compute_alignment_error(e, t, latent_dim=r)
    assert _close(al, G[f"{case}_aligned"]) and abs(err - float(G[f"{case}_error"])) <= TOL * abs(err)
    err0, al0 = ao.compute_alignment_error(e, t, latent_dim=None)
    assert _close(al0, G[f"{case}_aligned_signs"]) and abs(err0 - float(G[f"{case}_error_signs"])) <= TOL * abs(err0)
    assert _close(ao.align_latent_positions(e[:, 2:], t[:, 2:], r), G[f"{case}_latent"])


@pytest.mark.parametrize("case", ["p_d1", "p_d3", "p_d5", "p_d16"])
def test_procrustes_and_signs(case):
    e, t = G[f"{case}_est"], G[f"{case}_true"]
    al, R = ao.procrustes_alignment(e, t)
    assert _close(R, G[f"{case}_R"]) and _close(al, G[f"{case}_aligned"])
    assert _close(R @ R.T, np.eye(R.shape[0])) and np.linalg.det(R) > 0
    assert _close(ao.procrustes_alignment(e, t, scaling=True)[0], G[f"{case}_aligned_scaled"])
    assert np.array_equal(ao.align_signs(e, t, dim=0), G[f"{case}_signs_dim0"])
    assert np.array_equal(ao.align_signs(e, t, dim=1), G[f"{case}_signs_dim1"])


def test_latent_dim_required_for_temporal():
    with pytest.raises(ValueError):
        ao.compute_alignment_error(np.zeros((3, 2, 6)), np.zeros((3, 2, 6)))
