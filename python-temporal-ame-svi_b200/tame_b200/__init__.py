"""tame_b200 -- B200-native (sm_100a, FP64) variational inference for Temporal AME network models.

Drop-in surface of Alfieriek/Python-Temporal-AME-SVI for ONE path, the variational update loop:

    from tame_b200 import TemporalAMEModel, TemporalAMENaiveMFVI, TemporalAMEStructuredMFVI
    model = TemporalAMEModel(n_nodes=15, n_time=10); model.generate_data()
    vi = TemporalAMEStructuredMFVI(model, factorization="good", learning_rate=0.01)
    history = vi.fit(max_iter=100)

(`from src.models import ...` / `from src.inference import ...` work too when this directory is on sys.path.)
The sweep, the ELBO and the reconstruction error run as hand-written CUDA kernels behind the C ABI of
include/tame_b200.h; there is no CPU fallback.
"""
from .models import BaseAMEModel, StaticAMEModel, TemporalAMEModel
from .inference import (BaseVariationalInference, BaseTemporalVariationalInference, TemporalAMENaiveMFVI,
                        TemporalAMEStructuredMFVI, fit_batch)

__version__ = "0.1.0"
__all__ = ["BaseAMEModel", "StaticAMEModel", "TemporalAMEModel", "BaseVariationalInference",
           "BaseTemporalVariationalInference", "TemporalAMENaiveMFVI", "TemporalAMEStructuredMFVI", "fit_batch"]
