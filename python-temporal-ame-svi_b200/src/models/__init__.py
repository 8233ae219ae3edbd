"""reference: src/models/__init__.py:33-43"""
from tame_b200.models import BaseAMEModel, StaticAMEModel, TemporalAMEModel

__all__ = ["BaseAMEModel", "StaticAMEModel", "TemporalAMEModel"]
