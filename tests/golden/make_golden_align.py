#!/usr/bin/env python
"""Golden vectors for the step AFTER the fit -- src/utils/alignment.py of the UNMODIFIED reference
(/root/reference, read-only), run in float64.  Test infrastructure; only runs in the build container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_align.py

The reference module is loaded by file path (src/utils/__init__.py pulls in modules this container cannot
import).  One fixture, tests/golden/align.npz: for every case the inputs and what the reference returned.
"""
import importlib.util
import os
import sys

import numpy as np

REF = os.environ.get("TAME_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True

import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
torch.set_num_threads(1)

spec = importlib.util.spec_from_file_location("ref_alignment", os.path.join(REF, "src/utils/alignment.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

HERE = os.path.dirname(os.path.abspath(__file__))


def _rot(rng, k):
    q, _ = np.linalg.qr(rng.standard_normal((k, k)))
    return q


def make_states(rng, n, T, r, noise, mode):
    """X_true: AR-ish trajectories.  X_est: 'random' = unrelated draw; 'rotated' = per-time rotation/reflection of
    the U and V blocks, random row sign flips and noise (what a fit returns up to identifiability)."""
    d = 2 + 2 * r
    Xt = np.zeros((n, T, d))
    Xt[:, 0] = rng.standard_normal((n, d))
    for t in range(1, T):
        Xt[:, t] = 0.8 * Xt[:, t - 1] + 0.6 * rng.standard_normal((n, d))
    if mode == "random":
        return rng.standard_normal((n, T, d)), Xt
    Xe = Xt.copy()
    for t in range(T):
        Xe[:, t, 2:2 + r] = Xt[:, t, 2:2 + r] @ _rot(rng, r)
        Xe[:, t, 2 + r:] = Xt[:, t, 2 + r:] @ _rot(rng, r)
    Xe *= rng.choice([-1.0, 1.0], size=(n, T, 1))
    Xe += noise * rng.standard_normal(Xe.shape)
    return Xe, Xt


# name -> (n, T, r, noise, mode, align_each_time)
TEMPORAL = {
    "t_demo": (15, 10, 2, 0.2, "rotated", True),
    "t_r3": (40, 6, 3, 0.05, "rotated", True),
    "t_r1": (33, 5, 1, 0.3, "random", True),
    "t_r8": (64, 4, 8, 0.1, "rotated", True),
    "t_random": (25, 7, 4, 0.0, "random", True),
    "g_r2": (20, 7, 2, 0.2, "rotated", False),
    "g_r8": (50, 3, 8, 0.1, "random", False),
    "g_r5": (30, 4, 5, 0.1, "rotated", False),
}


def main():
    rng = np.random.default_rng(20251018)
    out = {"torch_version": np.array(torch.__version__), "temporal_cases": np.array(sorted(TEMPORAL))}
    for name, (n, T, r, noise, mode, each) in TEMPORAL.items():
        Xe, Xt = make_states(rng, n, T, r, noise, mode)
        te, tt = torch.from_numpy(Xe), torch.from_numpy(Xt)
        al = ref.align_temporal_states(te, tt, r, align_each_time=each)
        out[f"{name}_est"], out[f"{name}_true"], out[f"{name}_aligned"] = Xe, Xt, al.numpy().copy()
        out[f"{name}_r"], out[f"{name}_each"] = r, int(each)
        if each:
            err, al2 = ref.compute_alignment_error(te, tt, latent_dim=r, align=True)
            assert torch.equal(al, al2)
            out[f"{name}_error"] = err
            out[f"{name}_corr"] = ref.compute_correlation_after_alignment(te, tt, latent_dim=r)
            out[f"{name}_error_noalign"] = ref.compute_alignment_error(te, tt, latent_dim=r, align=False)[0]
        print(f"  {name}: n={n} T={T} r={r} each={each} mse after {((al - tt) ** 2).mean().item():.6f} "
              f"before {((te - tt) ** 2).mean().item():.6f}")
    # static (n, d) inputs: compute_alignment_error with and without latent_dim, the building blocks
    for name, (n, r) in {"s_r2": (30, 2), "s_r8": (45, 8)}.items():
        Xe, Xt = make_states(rng, n, 1, r, 0.1, "rotated")
        te, tt = torch.from_numpy(Xe[:, 0].copy()), torch.from_numpy(Xt[:, 0].copy())
        err, al = ref.compute_alignment_error(te, tt, latent_dim=r, align=True)
        err0, al0 = ref.compute_alignment_error(te, tt, latent_dim=None, align=True)
        out[f"{name}_est"], out[f"{name}_true"], out[f"{name}_r"] = te.numpy(), tt.numpy(), r
        out[f"{name}_aligned"], out[f"{name}_error"] = al.numpy().copy(), err
        out[f"{name}_aligned_signs"], out[f"{name}_error_signs"] = al0.numpy().copy(), err0
        out[f"{name}_latent"] = ref.align_latent_positions(te[:, 2:], tt[:, 2:], r).numpy().copy()
    for name, (n, d) in {"p_d3": (20, 3), "p_d5": (60, 5), "p_d16": (40, 16), "p_d1": (9, 1)}.items():
        A, B = rng.standard_normal((n, d)), rng.standard_normal((n, d))
        ta, tb = torch.from_numpy(A), torch.from_numpy(B)
        al, R = ref.procrustes_alignment(ta, tb, scaling=False)
        als, Rs = ref.procrustes_alignment(ta, tb, scaling=True)
        out[f"{name}_est"], out[f"{name}_true"] = A, B
        out[f"{name}_aligned"], out[f"{name}_R"] = al.numpy().copy(), R.numpy().copy()
        out[f"{name}_aligned_scaled"] = als.numpy().copy()
        out[f"{name}_signs_dim0"] = ref.align_signs(ta, tb, dim=0).numpy().copy()
        out[f"{name}_signs_dim1"] = ref.align_signs(ta, tb, dim=1).numpy().copy()
    np.savez_compressed(os.path.join(HERE, "align.npz"), **out)
    print("wrote", os.path.join(HERE, "align.npz"))


if __name__ == "__main__":
    main()
