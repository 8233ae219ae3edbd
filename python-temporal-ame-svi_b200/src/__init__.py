"""Import-compatibility shim: with python-temporal-ame-svi_b200/ on sys.path, reference-style imports
(`from src.models import TemporalAMEModel`, `from src.inference import TemporalAMEStructuredMFVI`) resolve to
the B200-native implementation in tame_b200 (reference: src/__init__.py:49-60)."""
__version__ = "0.1.0"
from .models import StaticAMEModel, TemporalAMEModel
from .inference import TemporalAMENaiveMFVI, TemporalAMEStructuredMFVI

__all__ = ["StaticAMEModel", "TemporalAMEModel", "TemporalAMENaiveMFVI", "TemporalAMEStructuredMFVI"]
