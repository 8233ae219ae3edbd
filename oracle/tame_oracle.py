"""CPU oracle for the Temporal-AME variational update loop.  TEST INFRASTRUCTURE ONLY.

Nothing in the product path (python-temporal-ame-svi_b200/) may import this module.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs use it, and only as the checker or the timed CPU baseline.

It is a NumPy float64 restatement of the reference algorithm
(Alfieriek/Python-Temporal-AME-SVI), one function per reference routine:

  model_constants     src/models/static_ame.py:96-127, src/models/base.py:146-196,
                      src/models/temporal_ame.py:129-145
  observation_terms   src/inference/structured_mf.py:289-326  (== naive_mf.py:284-376)
  update_node         src/inference/structured_mf.py:220-287, src/inference/naive_mf.py:207-282
  sweep               src/inference/structured_mf.py:211-218, src/inference/naive_mf.py:193-205
  elbo_parts          src/inference/structured_mf.py:115-209, src/inference/naive_mf.py:89-191
  compute_mean        src/models/static_ame.py:189-238
  reconstruction_mse  src/models/temporal_ame.py:255-291 (via src/inference/base.py:314-326)
  fit                 src/inference/base.py:127-208

PARITY PIN: the reference's own tests hold no numeric golden values for this path
(SURVEY.md section 8c), so the pin is the reference itself, run in float64 in the build
container by tests/golden/make_golden.py; tests/test_oracle.py checks every function
here against those fixtures (tests/golden/*.npz).

`sweep_blocked` is the schedule-preserving restructuring of SURVEY.md Appendix A
(running totals, static upper part, right-looking pushes, inline window) that the
CUDA path implements; it is checked against `sweep` (the literal schedule).
"""
from __future__ import annotations

import math

import numpy as np

NAIVE, GOOD, BAD = 0, 1, 2
MODE_OF = {"naive": NAIVE, "good": GOOD, "bad": BAD}
LOG_2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------------
# hyper-parameters
# ----------------------------------------------------------------------------------
def _equicorr(dim, corr, var):
    """src/models/base.py:146-160 (_generate_covariance_matrix)."""
    c = np.full((dim, dim), corr * var, dtype=np.float64)
    np.fill_diagonal(c, var)
    return c


def model_constants(n, T, r, ar_coefficient=0.8, rho_additive=0.5, rho_multiplicative=0.3,
                    rho_dyadic=0.5, process_noise_scale=0.1):
    """Hyper-parameters exactly as TemporalAMEModel builds them.

    R     static_ame.py:96-101 (variance 0.1, correlation rho_dyadic)
    Sigma static_ame.py:110-115, Psi static_ame.py:120-127 (two r x r equicorrelation blocks)
    Phi   temporal_ame.py:132, Q temporal_ame.py:139-145
    """
    d = 2 + 2 * r
    R = _equicorr(2, rho_dyadic, 0.1)
    Sigma = _equicorr(2, rho_additive, 1.0)
    Psi = np.zeros((2 * r, 2 * r))
    Psi[:r, :r] = _equicorr(r, rho_multiplicative, 1.0)
    Psi[r:, r:] = _equicorr(r, rho_multiplicative, 1.0)
    S0 = np.zeros((d, d))
    S0[:2, :2] = Sigma
    S0[2:, 2:] = Psi
    Phi = np.eye(d) * ar_coefficient
    Q = (1.0 - ar_coefficient ** 2) * S0
    Q = Q * process_noise_scale
    return derived_constants(dict(n=n, T=T, r=r, d=d, R=R, Sigma=Sigma, Psi=Psi, Phi=Phi, Q=Q))


def derived_constants(c):
    """Everything the sweep/ELBO derive from (R, Sigma, Psi, Phi, Q):
    structured_mf.py:229-237 (Q_inv, Sigma_0_inv), :128,:158,:180 (logdets)."""
    c = dict(c)
    d = c["d"]
    S0 = np.zeros((d, d))
    S0[:2, :2] = c["Sigma"]
    S0[2:, 2:] = c["Psi"]
    c["S0"] = S0
    c["R_inv"] = c.get("R_inv", np.linalg.inv(c["R"]))
    c["S0_inv"] = np.linalg.inv(S0)
    c["Q_inv"] = np.linalg.inv(c["Q"])
    c["logdet_R"] = np.linalg.slogdet(c["R"])[1]
    c["logdet_S0"] = np.linalg.slogdet(S0)[1]
    c["logdet_Q"] = np.linalg.slogdet(c["Q"])[1]
    return c


# ----------------------------------------------------------------------------------
# sweep (literal schedule)
# ----------------------------------------------------------------------------------
def observation_terms(Y, Xm, i, t, R_inv, r, row=None):
    """structured_mf.py:289-326.  P_obs = sum_{j!=i} J_j' R^-1 J_j, h_obs = sum J_j' R^-1 y_ij with
    J_j = [[1,0,V_j,0],[0,1,0,U_j]] built from the partners' CURRENT means at time t."""
    n = Y.shape[1]
    d = 2 + 2 * r
    mask = np.ones(n, dtype=bool)
    mask[i] = False
    U = Xm[mask, t, 2:2 + r]
    V = Xm[mask, t, 2 + r:]
    y = Y[i if row is None else row, mask, t, :]   # (n-1, 2); `row`: storage row of node i in a sharded Y
    m = n - 1
    J = np.zeros((m, 2, d))
    J[:, 0, 0] = 1.0
    J[:, 0, 2:2 + r] = V
    J[:, 1, 1] = 1.0
    J[:, 1, 2 + r:] = U
    RJ = np.einsum("ab,mbd->mad", R_inv, J)
    P = np.einsum("mad,mae->de", J, RJ)
    h = np.einsum("mad,ab,mb->d", J, R_inv, y)
    return P, h


def update_node(Y, Xm, Xc, i, c, lr, mode, row=None):
    """structured_mf.py:220-287 (good/bad) and naive_mf.py:207-282 (naive), in place."""
    T, r, d = c["T"], c["r"], c["d"]
    Phi, Q_inv, S0_inv, R_inv = c["Phi"], c["Q_inv"], c["S0_inv"], c["R_inv"]
    for t in range(T):
        P, h = observation_terms(Y, Xm, i, t, R_inv, r, row)
        if t == 0:
            P = P + S0_inv
        if t > 0:
            P = P + Q_inv
            h = h + Q_inv @ (Phi @ Xm[i, t - 1])
        if t < T - 1:
            P = P + Phi.T @ (Q_inv @ Phi)
            h = h + Phi.T @ (Q_inv @ Xm[i, t + 1])
        if mode == NAIVE:
            mu = np.linalg.solve(P, h)                          # naive_mf.py:268
            C = np.diag(1.0 / (np.diag(P) + 1e-8))              # naive_mf.py:271-274
        else:
            C = np.linalg.inv(P)                                # structured_mf.py:267
            if mode == BAD:
                C[:2, 2:] = 0.0                                 # :270-273
                C[2:, :2] = 0.0
            C = (C + C.T) / 2                                   # :276
            C = C + np.eye(d) * 1e-6                            # :277
            mu = C @ h                                          # :279
        Xm[i, t] = lr * mu + (1 - lr) * Xm[i, t]                # :282-287
        Xc[i, t] = lr * C + (1 - lr) * Xc[i, t]


def sweep(Y, Xm, Xc, c, lr, mode):
    """structured_mf.py:211-218: Gauss-Seidel over nodes 0..n-1, times 0..T-1 inside."""
    for i in range(c["n"]):
        update_node(Y, Xm, Xc, i, c, lr, mode)


def sweep_sharded(Y_local, Xm, Xc, c, lr, mode, world, rank, panel, broadcast):
    """The multi-GPU schedule on one rank: nodes in order; the owner of a panel updates its nodes from ITS rows of Y
    (Y_local in storage order) and the replicated means, then `broadcast(root, array)` ships the panel's new means.
    X_cov rows of foreign nodes are left untouched (tame_gather_state collects them).  Same schedule as `sweep`."""
    n = c["n"]
    for lo in range(0, n, panel):
        hi = min(n, lo + panel)
        owner = (lo // panel) % world
        if owner == rank:
            for i in range(lo, hi):
                b = i // panel
                update_node(Y_local, Xm, Xc, i, c, lr, mode, row=(b // world) * panel + (i - b * panel))
        blk = np.ascontiguousarray(Xm[lo:hi])
        broadcast(owner, blk)
        Xm[lo:hi] = blk


# ----------------------------------------------------------------------------------
# sweep, blocked (the restructuring the CUDA path uses)
# ----------------------------------------------------------------------------------
def _moment_totals(M, r):
    U, V = M[..., :r], M[..., r:]
    return dict(sU=U.sum(0), sV=V.sum(0), SUU=np.einsum("jta,jtb->tab", U, U),
                SVV=np.einsum("jta,jtb->tab", V, V), SVU=np.einsum("jta,jtb->tab", V, U))


def sweep_blocked(Y, Xm, Xc, c, lr, mode, block=4):
    """Same schedule as `sweep`, organised the way the GPU kernels are:

      hab[i,t]   = sum_{j!=i} (w0, w1)                     constant in the sweep
      H[i,t]     = sum_{j>i} w0*V_j^old , w1*U_j^old       static upper part (sweep start)
      push(blk)  : H[k,t] += sum_{j in blk} w*M_j^new      for rows k >= (blk+2)*block
      inline     : sum over j in [ (blk(i)-1)*block , i )  with the new means
      totals     : running sums of U,V,UU',VV',VU' with node i's own term removed/added.
    """
    n, T, r, d = c["n"], c["T"], c["r"], c["d"]
    p, q = c["R_inv"][0, 0], c["R_inv"][0, 1]
    Phi, Q_inv, S0_inv = c["Phi"], c["Q_inv"], c["S0_inv"]
    W0 = p * Y[..., 0] + q * Y[..., 1]          # (n,n,T)
    W1 = q * Y[..., 0] + p * Y[..., 1]
    off = ~np.eye(n, dtype=bool)
    hab = np.stack([(W0 * off[:, :, None]).sum(1), (W1 * off[:, :, None]).sum(1)], -1)   # (n,T,2)
    up = np.triu(np.ones((n, n), dtype=bool), 1)
    M_old = Xm[:, :, 2:].copy()
    HU = np.einsum("ijt,jta->ita", W0 * up[:, :, None], M_old[..., r:])
    HV = np.einsum("ijt,jta->ita", W1 * up[:, :, None], M_old[..., :r])
    tot = _moment_totals(M_old, r)
    QiPhi = Q_inv @ Phi
    PhiTQi = Phi.T @ Q_inv
    PhiTQiPhi = Phi.T @ (Q_inv @ Phi)
    nblk = (n + block - 1) // block
    for blk in range(nblk):
        lo, hi = blk * block, min(n, (blk + 1) * block)
        for i in range(lo, hi):
            wlo = max(0, (blk - 1) * block)       # inline window start
            for t in range(T):
                U_i, V_i = Xm[i, t, 2:2 + r].copy(), Xm[i, t, 2 + r:].copy()
                sU = tot["sU"][t] - U_i
                sV = tot["sV"][t] - V_i
                SUU = tot["SUU"][t] - np.outer(U_i, U_i)
                SVV = tot["SVV"][t] - np.outer(V_i, V_i)
                SVU = tot["SVU"][t] - np.outer(V_i, U_i)
                P = np.zeros((d, d))
                P[0, 0] = P[1, 1] = p * (n - 1)
                P[0, 1] = P[1, 0] = q * (n - 1)
                P[0, 2:2 + r] = P[2:2 + r, 0] = p * sV
                P[0, 2 + r:] = P[2 + r:, 0] = q * sU
                P[1, 2:2 + r] = P[2:2 + r, 1] = q * sV
                P[1, 2 + r:] = P[2 + r:, 1] = p * sU
                P[2:2 + r, 2:2 + r] = p * SVV
                P[2 + r:, 2 + r:] = p * SUU
                P[2:2 + r, 2 + r:] = q * SVU
                P[2 + r:, 2:2 + r] = q * SVU.T
                h = np.zeros(d)
                h[0:2] = hab[i, t]
                js = np.arange(wlo, i)
                h[2:2 + r] = HU[i, t] + W0[i, js, t] @ Xm[js, t, 2 + r:]
                h[2 + r:] = HV[i, t] + W1[i, js, t] @ Xm[js, t, 2:2 + r]
                if t == 0:
                    P += S0_inv
                if t > 0:
                    P += Q_inv
                    h += QiPhi @ Xm[i, t - 1]
                if t < T - 1:
                    P += PhiTQiPhi
                    h += PhiTQi @ Xm[i, t + 1]
                if mode == NAIVE:
                    mu = np.linalg.solve(P, h)
                    C = np.diag(1.0 / (np.diag(P) + 1e-8))
                else:
                    C = np.linalg.inv(P)
                    if mode == BAD:
                        C[:2, 2:] = 0.0
                        C[2:, :2] = 0.0
                    C = (C + C.T) / 2 + np.eye(d) * 1e-6
                    mu = C @ h
                Xm[i, t] = lr * mu + (1 - lr) * Xm[i, t]
                Xc[i, t] = lr * C + (1 - lr) * Xc[i, t]
                Un, Vn = Xm[i, t, 2:2 + r], Xm[i, t, 2 + r:]
                tot["sU"][t] = sU + Un
                tot["sV"][t] = sV + Vn
                tot["SUU"][t] = SUU + np.outer(Un, Un)
                tot["SVV"][t] = SVV + np.outer(Vn, Vn)
                tot["SVU"][t] = SVU + np.outer(Vn, Un)
        # right-looking push of the finished block to rows two blocks further on
        k0 = (blk + 2) * block
        if k0 < n:
            HU[k0:] += np.einsum("kjt,jta->kta", W0[k0:, lo:hi], Xm[lo:hi, :, 2 + r:])
            HV[k0:] += np.einsum("kjt,jta->kta", W1[k0:, lo:hi], Xm[lo:hi, :, 2:2 + r])


# ----------------------------------------------------------------------------------
# ELBO + reconstruction error
# ----------------------------------------------------------------------------------
def compute_mean(A, M, r):
    """static_ame.py:189-238: mu[i,j,0]=a_i+b_j+U_i.V_j ; mu[i,j,1]=a_j+b_i+U_j.V_i"""
    a, b = A[:, 0], A[:, 1]
    U, V = M[:, :r], M[:, r:]
    add = a[:, None] + b[None, :]
    mul = U @ V.T
    return np.stack([add + mul, add.T + mul.T], -1)


def elbo_parts(Y, Xm, Xc, c, mode):
    """(LL, LP0, LPT, H) of structured_mf.py:124-209 / naive_mf.py:114-191."""
    n, T, r, d = c["n"], c["T"], c["r"], c["d"]
    R_inv = c["R_inv"]
    iu = np.triu_indices(n, 1)
    ll = 0.0
    tr_Rinv = np.trace(R_inv)
    for t in range(T):
        mu = compute_mean(Xm[:, t, :2], Xm[:, t, 2:], r)
        res = (Y[:, :, t] - mu)[iu]                                   # (pairs, 2), i<j only
        quad = np.einsum("ma,ab,mb->m", res, R_inv, res)
        if mode == NAIVE:
            corr = 0.0                                                # naive_mf.py:128-130
        else:
            tr = np.trace(Xc[:, t], axis1=1, axis2=2)
            corr = 0.1 * (tr[iu[0]] + tr[iu[1]]) * tr_Rinv / d        # structured_mf.py:142-144
        ll += np.sum(-0.5 * (c["logdet_R"] + quad + corr + 2 * LOG_2PI))
    lp0, lpt, ent = elbo_state_parts(Xm, Xc, c)
    return float(ll), lp0, lpt, ent


def elbo_state_parts(Xm, Xc, c):
    """(LP0, LPT, H): the ELBO terms that only read the variational state -- structured_mf.py:148-209 /
    naive_mf.py:134-191.  O(n T d^3), so it also serves at sizes where Y does not fit on the host."""
    T, d = c["T"], c["d"]
    quad0 = np.einsum("ia,ab,ib->i", Xm[:, 0], c["S0_inv"], Xm[:, 0])
    tr0 = np.einsum("ab,iba->i", c["S0_inv"], Xc[:, 0])
    lp0 = np.sum(-0.5 * (c["logdet_S0"] + quad0 + tr0 + d * LOG_2PI))
    lpt = 0.0
    if T > 1:
        resid = Xm[:, 1:] - Xm[:, :-1] @ c["Phi"].T
        quadt = np.einsum("ita,ab,itb->it", resid, c["Q_inv"], resid)
        trt = np.einsum("ab,itba->it", c["Q_inv"], Xc[:, 1:])
        lpt = np.sum(-0.5 * (c["logdet_Q"] + quadt + trt + d * LOG_2PI))
    logdet = np.linalg.slogdet(Xc)[1]
    ent = np.sum(0.5 * (d * (1 + LOG_2PI) + logdet))
    return float(lp0), float(lpt), float(ent)


def elbo(Y, Xm, Xc, c, mode):
    return float(sum(elbo_parts(Y, Xm, Xc, c, mode)))


def reconstruction_mse(Y, Xm, c):
    """temporal_ame.py:255-291: sum over t, i!=j and BOTH components, divided by n(n-1)T."""
    n, T, r = c["n"], c["T"], c["r"]
    off = ~np.eye(n, dtype=bool)
    tot = 0.0
    for t in range(T):
        mu = compute_mean(Xm[:, t, :2], Xm[:, t, 2:], r)
        tot += float((((Y[:, :, t] - mu) ** 2) * off[:, :, None]).sum())
    return tot / (n * (n - 1) * T)


# ----------------------------------------------------------------------------------
# fit loop
# ----------------------------------------------------------------------------------
def fit(Y, Xm, Xc, c, lr, mode, max_iter=100, tolerance=1e-4, blocked=False, block=4):
    """base.py:127-208: sweep -> ELBO -> MSE, early stop after 3 consecutive iterations with
    |dELBO|/(|ELBO_prev|+1e-8) < tolerance (never tested on the first iteration)."""
    elbos, mses = [], []
    patience = 0
    prev = -np.inf
    for it in range(max_iter):
        if blocked:
            sweep_blocked(Y, Xm, Xc, c, lr, mode, block=block)
        else:
            sweep(Y, Xm, Xc, c, lr, mode)
        e = elbo(Y, Xm, Xc, c, mode)
        elbos.append(e)
        mses.append(reconstruction_mse(Y, Xm, c))
        converged = False
        if it > 0:
            rel = abs(e - prev) / (abs(prev) + 1e-8)
            patience = patience + 1 if rel < tolerance else 0
            converged = patience >= 3
        prev = e
        if converged:
            break
    return np.array(elbos), np.array(mses)


# ----------------------------------------------------------------------------------
# fast variant for CPU-baseline timing at sizes the literal form cannot reach
# ----------------------------------------------------------------------------------
def sweep_fast(Y, Xm, Xc, c, lr, mode):
    """Literal Gauss-Seidel order, partner sums via BLAS (one (i,t) cell per step, partners
    vectorised).  Used as the timed 'port' CPU baseline; checked against `sweep`."""
    n, T, r, d = c["n"], c["T"], c["r"], c["d"]
    p, q = c["R_inv"][0, 0], c["R_inv"][0, 1]
    Phi, Q_inv, S0_inv = c["Phi"], c["Q_inv"], c["S0_inv"]
    QiPhi = Q_inv @ Phi
    PhiTQi = Phi.T @ Q_inv
    PhiTQiPhi = Phi.T @ (Q_inv @ Phi)
    eye = np.eye(d)
    for i in range(n):
        Yi = Y[i]                                   # (n,T,2)
        w0 = p * Yi[..., 0] + q * Yi[..., 1]        # (n,T)
        w1 = q * Yi[..., 0] + p * Yi[..., 1]
        w0[i] = 0.0
        w1[i] = 0.0
        ha, hb = w0.sum(0), w1.sum(0)
        for t in range(T):
            Mt = Xm[:, t, 2:]
            U, V = Mt[:, :r], Mt[:, r:]
            Ui, Vi = U[i].copy(), V[i].copy()
            sU, sV = U.sum(0) - Ui, V.sum(0) - Vi
            SUU = U.T @ U - np.outer(Ui, Ui)
            SVV = V.T @ V - np.outer(Vi, Vi)
            SVU = V.T @ U - np.outer(Vi, Ui)
            P = np.zeros((d, d))
            P[0, 0] = P[1, 1] = p * (n - 1)
            P[0, 1] = P[1, 0] = q * (n - 1)
            P[0, 2:2 + r] = P[2:2 + r, 0] = p * sV
            P[0, 2 + r:] = P[2 + r:, 0] = q * sU
            P[1, 2:2 + r] = P[2:2 + r, 1] = q * sV
            P[1, 2 + r:] = P[2 + r:, 1] = p * sU
            P[2:2 + r, 2:2 + r] = p * SVV
            P[2 + r:, 2 + r:] = p * SUU
            P[2:2 + r, 2 + r:] = q * SVU
            P[2 + r:, 2:2 + r] = q * SVU.T
            h = np.empty(d)
            h[0], h[1] = ha[t], hb[t]
            h[2:2 + r] = w0[:, t] @ V
            h[2 + r:] = w1[:, t] @ U
            if t == 0:
                P += S0_inv
            if t > 0:
                P += Q_inv
                h += QiPhi @ Xm[i, t - 1]
            if t < T - 1:
                P += PhiTQiPhi
                h += PhiTQi @ Xm[i, t + 1]
            if mode == NAIVE:
                mu = np.linalg.solve(P, h)
                C = np.diag(1.0 / (np.diag(P) + 1e-8))
            else:
                C = np.linalg.inv(P)
                if mode == BAD:
                    C[:2, 2:] = 0.0
                    C[2:, :2] = 0.0
                C = (C + C.T) / 2 + eye * 1e-6
                mu = C @ h
            Xm[i, t] = lr * mu + (1 - lr) * Xm[i, t]
            Xc[i, t] = lr * C + (1 - lr) * Xc[i, t]


def elbo_mse_fast(Y, Xm, Xc, c, mode):
    """ELBO + MSE in one vectorised pass (same formulas as elbo_parts/reconstruction_mse)."""
    return elbo(Y, Xm, Xc, c, mode), reconstruction_mse(Y, Xm, c)
