"""reference: src/inference/__init__.py:45-56"""
from tame_b200.inference import (BaseVariationalInference, BaseTemporalVariationalInference, TemporalAMENaiveMFVI,
                                 TemporalAMEStructuredMFVI)

__all__ = ["BaseVariationalInference", "BaseTemporalVariationalInference", "TemporalAMENaiveMFVI",
           "TemporalAMEStructuredMFVI"]
