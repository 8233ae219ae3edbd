#!/usr/bin/env python
"""Golden vectors for the contribution / U'V diagnostics from the UNMODIFIED reference:
src/utils/diagnostics.py (loaded by file path) and compute_uv_correlation_over_time of
experiments/multiplicative_strength_comparison.py:46-89 (that script imports matplotlib, so the function's source text is
executed on its own -- nothing is copied into the repository).  float64; build container only.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_diag.py
"""
import ast
import importlib.util
import os
import sys

import numpy as np

REF = os.environ.get("TAME_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True

import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
torch.set_num_threads(1)

spec = importlib.util.spec_from_file_location("ref_diagnostics", os.path.join(REF, "src/utils/diagnostics.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

# pull the one function out of the experiment script without importing the script
src = open(os.path.join(REF, "experiments/multiplicative_strength_comparison.py")).read()
fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "compute_uv_correlation_over_time"][0]
ns = {"torch": torch}
exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref_uv_corr", "exec"), ns)
uv_corr_over_time = ns["compute_uv_correlation_over_time"]

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"d_demo": (15, 10, 2), "d_r3": (40, 6, 3), "d_r1": (33, 5, 1), "d_r8": (64, 4, 8), "d_r5": (27, 3, 5)}


def main():
    rng = np.random.default_rng(20251019)
    out = {"cases": np.array(sorted(CASES)), "torch_version": np.array(torch.__version__)}
    for name, (n, T, r) in CASES.items():
        d = 2 + 2 * r
        Xt = rng.standard_normal((n, T, d)) * np.array([1.0, 0.7] + [0.8] * (2 * r))
        Xe = 0.8 * Xt + 0.4 * rng.standard_normal((n, T, d)) + 0.05
        te, tt = torch.from_numpy(Xe), torch.from_numpy(Xt)
        for tag, excl in (("excl", True), ("incl", False)):
            add, mul = ref.compute_temporal_contributions(te, r, exclude_diagonal=excl)
            out[f"{name}_add_{tag}"], out[f"{name}_mul_{tag}"] = add.numpy().copy(), mul.numpy().copy()
        out[f"{name}_est"], out[f"{name}_true"], out[f"{name}_r"] = Xe, Xt, r
        out[f"{name}_ratio0"] = ref.compute_contribution_ratio(te[:, 0, :2], te[:, 0, 2:])
        out[f"{name}_state_mse"] = ref.compute_state_prediction_error(tt, te)
        out[f"{name}_uvcorr_t"] = uv_corr_over_time(te, tt, r).numpy().copy()
        out[f"{name}_uvcorr0"] = ref.compute_uv_product_correlation(te[:, 0, 2:], tt[:, 0, 2:], r)
        print(f"  {name}: add[0]={float(out[f'{name}_add_excl'][0]):.6f} mul[0]={float(out[f'{name}_mul_excl'][0]):.6f} "
              f"uvcorr[0]={float(out[f'{name}_uvcorr_t'][0]):.6f}")
    np.savez_compressed(os.path.join(HERE, "diag.npz"), **out)
    print("wrote", os.path.join(HERE, "diag.npz"))


if __name__ == "__main__":
    main()
