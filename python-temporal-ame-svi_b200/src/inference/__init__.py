"""reference: src/inference/__init__.py:45-56 (+ fit_batch, additive: many independent fits in one call)"""
from tame_b200.inference import (BaseVariationalInference, BaseTemporalVariationalInference, TemporalAMENaiveMFVI,
                                 TemporalAMEStructuredMFVI, fit_batch)

__all__ = ["BaseVariationalInference", "BaseTemporalVariationalInference", "TemporalAMENaiveMFVI",
           "TemporalAMEStructuredMFVI", "fit_batch"]
