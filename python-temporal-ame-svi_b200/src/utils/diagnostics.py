"""`from src.utils.diagnostics import ...` -> tame_b200.diagnostics (reference: src/utils/diagnostics.py)."""
from tame_b200.diagnostics import (compute_additive_contribution, compute_multiplicative_contribution,
                                   compute_temporal_contributions, compute_contribution_ratio, compute_state_prediction_error,
                                   compute_uv_product_correlation, compute_uv_correlation_over_time)

__all__ = ["compute_additive_contribution", "compute_multiplicative_contribution", "compute_temporal_contributions",
           "compute_contribution_ratio", "compute_state_prediction_error", "compute_uv_product_correlation",
           "compute_uv_correlation_over_time"]
