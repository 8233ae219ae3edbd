"""Import-compatibility shim for the reference's `src.utils` (reference: src/utils/__init__.py:30-50): the alignment
functions and the contribution / U'V diagnostics, which run on the GPU (tame_b200.alignment, tame_b200.diagnostics).
The printing helpers and the metrics module of the reference are outside the accelerated path and are not provided
here -- keep importing them from the reference checkout."""
from .alignment import (procrustes_alignment, align_signs, align_latent_positions, align_temporal_states,
                        compute_alignment_error, compute_correlation_after_alignment)
from .diagnostics import (compute_additive_contribution, compute_multiplicative_contribution, compute_temporal_contributions,
                          compute_contribution_ratio, compute_state_prediction_error, compute_uv_product_correlation,
                          compute_uv_correlation_over_time)

__all__ = ["procrustes_alignment", "align_signs", "align_latent_positions", "align_temporal_states",
           "compute_alignment_error", "compute_correlation_after_alignment",
           "compute_additive_contribution", "compute_multiplicative_contribution", "compute_temporal_contributions",
           "compute_contribution_ratio", "compute_state_prediction_error", "compute_uv_product_correlation",
           "compute_uv_correlation_over_time"]
