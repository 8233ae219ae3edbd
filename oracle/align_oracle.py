"""CPU oracle for the alignment step that follows every fit in the reference's drivers -- a NumPy restatement of
/root/reference/src/utils/alignment.py.

TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product path (tame_b200.alignment calls the CUDA library and
fails loudly without it).

Parity pin: tests/golden/align.npz holds outputs of the unmodified reference module run in float64 in the build
container (tests/golden/make_golden_align.py); tests/test_oracle_align.py checks every function below against it.

Reference quirks kept on purpose:
  * procrustes_alignment forms M = X_true' X_est, takes U S Vt = svd(M) and rotates with R = U Vt (alignment.py:76-92);
    R is the orthogonal polar factor of M.  When det(R) < 0 the LAST row of Vt (smallest singular value) is negated.
  * align_signs(dim = last) loops over ROWS and flips a whole row when ||-x - y|| < ||x - y|| (alignment.py:138-146);
    align_temporal_states uses it for the additive pair and, after the rotation, for the U and V rows separately.
  * the global mode rotates the whole 2r-dimensional multiplicative block with one 2r x 2r Procrustes of the temporal
    means (alignment.py:286-311), not U and V separately.
"""
import numpy as np


def procrustes_rotation(X_est, X_true):
    """alignment.py:76-88: R = U Vt of svd(X_true' X_est), last row of Vt negated if det(R) < 0."""
    M = X_true.T @ X_est
    U, _, Vt = np.linalg.svd(M)
    R = U @ Vt
    if np.linalg.det(R) < 0:
        Vt = Vt.copy()
        Vt[-1, :] *= -1
        R = U @ Vt
    return R


def procrustes_alignment(X_est, X_true, scaling=False):
    """alignment.py:31-103."""
    R = procrustes_rotation(X_est, X_true)
    Xa = X_est @ R
    if scaling:
        num = np.trace(X_true.T @ Xa)
        den = np.trace(Xa.T @ Xa)
        if den > 1e-10:
            Xa = Xa * (num / den)
    return Xa, R


def align_signs_rows(X_est, X_true):
    """alignment.py:138-146 (dim == last): per row, flip when the flipped row is strictly closer."""
    pos = np.sqrt(((X_est - X_true) ** 2).sum(-1))
    neg = np.sqrt(((-X_est - X_true) ** 2).sum(-1))
    return np.where((neg < pos)[..., None], -X_est, X_est)


def align_signs(X_est, X_true, dim=-1):
    """alignment.py:106-166: dim == last -> rows; otherwise whole slices along `dim`."""
    if dim == -1 or dim == X_est.ndim - 1:
        flat_e = X_est.reshape(X_est.shape[0], -1) if X_est.ndim > 2 else X_est
        if X_est.ndim > 2:          # the reference indexes X[i] (everything behind the first axis) in this branch
            flat_t = X_true.reshape(X_true.shape[0], -1)
            return align_signs_rows(flat_e, flat_t).reshape(X_est.shape)
        if X_est.ndim == 1:         # X[i] is a scalar
            return align_signs_rows(X_est[:, None], X_true[:, None])[:, 0]
        return align_signs_rows(X_est, X_true)
    e = np.moveaxis(X_est, dim, 0)
    t = np.moveaxis(X_true, dim, 0)
    out = align_signs_rows(e.reshape(e.shape[0], -1), t.reshape(t.shape[0], -1)).reshape(e.shape)
    return np.moveaxis(out, 0, dim)


def align_latent_positions(M_est, M_true, r):
    """alignment.py:169-221: U and V separately: Procrustes, then row signs."""
    U = align_signs_rows(procrustes_alignment(M_est[:, :r], M_true[:, :r])[0], M_true[:, :r])
    V = align_signs_rows(procrustes_alignment(M_est[:, r:], M_true[:, r:])[0], M_true[:, r:])
    return np.concatenate([U, V], axis=1)


def align_temporal_states(X_est, X_true, r, align_each_time=True):
    """alignment.py:224-313."""
    n, T, d = X_est.shape
    out = X_est.copy()
    if align_each_time:
        for t in range(T):
            out[:, t, :2] = align_signs_rows(X_est[:, t, :2], X_true[:, t, :2])
            out[:, t, 2:] = align_latent_positions(X_est[:, t, 2:], X_true[:, t, 2:], r)
        return out
    me, mt = X_est.mean(axis=1), X_true.mean(axis=1)
    R = procrustes_rotation(me[:, 2:], mt[:, 2:])
    for t in range(T):
        out[:, t, :2] = align_signs_rows(X_est[:, t, :2], X_true[:, t, :2])
        out[:, t, 2:] = align_signs_rows(X_est[:, t, 2:] @ R, X_true[:, t, 2:])
    return out


def compute_alignment_error(X_est, X_true, latent_dim=None, align=True):
    """alignment.py:316-385: (mean squared error after alignment, aligned states)."""
    Xa = X_est
    if align:
        if X_est.ndim == 3:
            if latent_dim is None:
                raise ValueError("latent_dim must be provided for temporal alignment")
            Xa = align_temporal_states(X_est, X_true, latent_dim)
        elif X_est.ndim == 2:
            if latent_dim is not None:
                Xa = np.concatenate([align_signs_rows(X_est[:, :2], X_true[:, :2]),
                                     align_latent_positions(X_est[:, 2:], X_true[:, 2:], latent_dim)], axis=1)
            else:
                Xa = align_signs_rows(X_est, X_true)
    return float(((Xa - X_true) ** 2).mean()), Xa


def compute_correlation_after_alignment(X_est, X_true, latent_dim=None):
    """alignment.py:388-435."""
    _, Xa = compute_alignment_error(X_est, X_true, latent_dim, align=True)
    a = Xa.ravel() - Xa.mean()
    b = X_true.ravel() - X_true.mean()
    den = np.sqrt((a * a).sum() * (b * b).sum())
    if den < 1e-10:
        return 0.0
    return float((a * b).sum() / den)
