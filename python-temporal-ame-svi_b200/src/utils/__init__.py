"""Import-compatibility shim for the reference's `src.utils` (reference: src/utils/__init__.py:43-50): the alignment
functions, which run on the GPU (tame_b200.alignment).  The diagnostics and metrics modules of the reference are
outside the accelerated path and are not provided here -- keep importing them from the reference checkout."""
from .alignment import (procrustes_alignment, align_signs, align_latent_positions, align_temporal_states,
                        compute_alignment_error, compute_correlation_after_alignment)

__all__ = ["procrustes_alignment", "align_signs", "align_latent_positions", "align_temporal_states",
           "compute_alignment_error", "compute_correlation_after_alignment"]
