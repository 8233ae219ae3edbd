"""CPU oracle for the contribution / U'V-product diagnostics -- NumPy restatement of
/root/reference/src/utils/diagnostics.py:82-273,528-561 and experiments/multiplicative_strength_comparison.py:46-89.

TEST INFRASTRUCTURE ONLY (tests/ and bench.py's CPU leg); the product path is tame_b200.diagnostics -> libtame_b200.
Parity pin: tests/golden/diag.npz (outputs of the unmodified reference, tests/golden/make_golden_diag.py).
The restatement is literal (n x n products per time step), unlike the device kernels which use the Gram identities
sum_ij (U_i.V_j)^2 = <U'U, V'V> etc.
"""
import numpy as np


def additive_contribution(A, exclude_diagonal=True):
    """diagnostics.py:82-122: mean over pairs of (a_i + b_j)^2."""
    n = A.shape[0]
    add = A[:, 0][:, None] + A[:, 1][None, :]
    if exclude_diagonal:
        return float(((add ** 2) * (1 - np.eye(n))).sum() / (n * (n - 1)))
    return float((add ** 2).sum() / (n * n))


def multiplicative_contribution(M, exclude_diagonal=True):
    """diagnostics.py:125-167: mean over pairs of (U_i . V_j)^2."""
    n, r = M.shape[0], M.shape[1] // 2
    mul = M[:, :r] @ M[:, r:].T
    if exclude_diagonal:
        return float(((mul ** 2) * (1 - np.eye(n))).sum() / (n * (n - 1)))
    return float((mul ** 2).sum() / (n * n))


def temporal_contributions(X, r, exclude_diagonal=True):
    """diagnostics.py:170-217."""
    T = X.shape[1]
    add = np.array([additive_contribution(X[:, t, :2], exclude_diagonal) for t in range(T)])
    mul = np.array([multiplicative_contribution(X[:, t, 2:], exclude_diagonal) for t in range(T)])
    return add, mul


def contribution_ratio(A, M):
    """diagnostics.py:220-251."""
    va, vm = additive_contribution(A), multiplicative_contribution(M)
    return float("inf") if vm < 1e-10 else float(np.sqrt(va / vm))


def state_prediction_error(X_true, X_pred):
    """diagnostics.py:254-273."""
    return float(((X_true - X_pred) ** 2).mean())


def uv_product_correlation(M_est, M_true, r):
    """diagnostics.py:528-561: Pearson correlation of the flattened n x n products (diagonal included)."""
    pe = (M_est[:, :r] @ M_est[:, r:].T).ravel()
    pt = (M_true[:, :r] @ M_true[:, r:].T).ravel()
    return float(np.corrcoef(np.stack([pt, pe]))[0, 1])


def uv_correlation_over_time(X_est, X_true, r):
    """multiplicative_strength_comparison.py:46-89."""
    return np.array([uv_product_correlation(X_est[:, t, 2:], X_true[:, t, 2:], r) for t in range(X_est.shape[1])])
