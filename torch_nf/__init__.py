"""Drop-in import name: ``import torch_nf.bijectors`` etc. resolve to the B200-native implementation.

The reference package is called ``torch_nf`` (setup.py:5-15) and its users and tests import
``torch_nf.bijectors``, ``torch_nf.density_estimator``, ``torch_nf.conditional_density_estimator`` and
``torch_nf.error_formatters`` (and ``torch_nf.exponential_families``, ``torch_nf.lfi`` - the latter is missing from the
reference repository but imported by ``scripts/lfi_mat.py:5``); ``ConditionalDensityEstimator`` accepts only the exact ``NormFlow`` type
(conditional_density_estimator.py:48), so the names must resolve to the SAME module objects as
``torch_nf_b200.*`` -- they are aliased in ``sys.modules``, not re-exported copies.
"""
import sys

import torch_nf_b200
from torch_nf_b200 import bijectors, conditional_density_estimator, density_estimator, error_formatters
from torch_nf_b200 import exponential_families, lfi
from torch_nf_b200 import set_conditioner_precision  # noqa: F401

for _name, _mod in (("bijectors", bijectors), ("density_estimator", density_estimator),
                    ("conditional_density_estimator", conditional_density_estimator),
                    ("error_formatters", error_formatters), ("exponential_families", exponential_families),
                    ("lfi", lfi)):
    sys.modules[__name__ + "." + _name] = _mod
__version__ = getattr(torch_nf_b200, "__version__", "0")
