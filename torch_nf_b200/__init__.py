"""torch_nf_b200: B200-native bijector-chain hot path behind the torch_nf API.

Drop-in for ``torch_nf.bijectors`` / ``torch_nf.density_estimator.NormFlow`` /
``torch_nf.conditional_density_estimator.ConditionalDensityEstimator``:

    import torch_nf_b200.density_estimator as de
    nf = de.NormFlow(64, False, "coupling", 4, 2, 256)
    z, log_q_z = nf(N=1 << 20)
    log_p = nf.log_prob(z)

All arithmetic runs in hand-written sm_100a CUDA kernels reached through the
C ABI in ``include/tnf.h`` (``torch_nf_b200/_C.so``).  There is no CPU path.
"""
from . import config  # noqa: F401
from .config import set_conditioner_precision, conditioner_precision, set_training_backward  # noqa: F401

__all__ = ["bijectors", "density_estimator", "conditional_density_estimator", "error_formatters", "config",
           "dist", "set_conditioner_precision", "conditioner_precision", "set_training_backward"]
