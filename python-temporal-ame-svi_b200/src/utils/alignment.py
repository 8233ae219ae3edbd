"""`from src.utils.alignment import ...` -> tame_b200.alignment (reference: src/utils/alignment.py)."""
from tame_b200.alignment import (procrustes_alignment, align_signs, align_latent_positions, align_temporal_states,
                                 compute_alignment_error, compute_correlation_after_alignment)

__all__ = ["procrustes_alignment", "align_signs", "align_latent_positions", "align_temporal_states",
           "compute_alignment_error", "compute_correlation_after_alignment"]
