"""Conditional density estimator: hyper-network ``x -> params`` in front of a
``NormFlow`` (reference torch_nf/conditional_density_estimator.py:10-104).

The boundary caller of the hot path: ``param_net`` stays a plain
``torch.nn.Sequential`` of ``Linear`` + ``Tanh`` (so ``state_dict`` keeps the
reference's keys ``linear1``, ``tanh1``, ``linear2``, ``relu2`` ...), and the
flow it parameterises runs on the CUDA kernels with one weight row per
context (regime B).
"""
from collections import OrderedDict

import torch

from . import density_estimator as de
from .error_formatters import format_type_err_msg


class ConditionalDensityEstimator(torch.nn.Module):
    def __init__(self, density_estimator, D_x, hidden_layers, dropout=False):
        super().__init__()
        self.density_estimator = density_estimator
        self.D_x = D_x
        self.D_params = density_estimator.D_params
        self.hidden_layers = hidden_layers
        self.dropout = dropout

        widths = [D_x] + list(self.hidden_layers)
        layers = []
        for i in range(1, len(widths)):
            layers.append(("linear%d" % i, torch.nn.Linear(widths[i - 1], widths[i])))
            # the reference names the first activation tanh1 and the later ones relu<i>; all are Tanh (:20-31)
            layers.append((("tanh%d" if i == 1 else "relu%d") % i, torch.nn.Tanh()))
            if self.dropout:
                layers.append(("dropout%d" % i, torch.nn.Dropout()))
        layers.append(("linear%d" % len(widths), torch.nn.Linear(widths[-1], self.D_params)))
        self.param_net = torch.nn.Sequential(OrderedDict(layers))

    @property
    def density_estimator(self):
        return self._density_estimator

    @density_estimator.setter
    def density_estimator(self, val):
        # exact type, as in the reference (:48): subclasses are rejected
        if type(val) is not de.NormFlow:
            raise TypeError(format_type_err_msg(self, "density_estimator", val, de.DensityEstimator))
        self._density_estimator = val

    @property
    def D_x(self):
        return self._D_x

    @D_x.setter
    def D_x(self, val):
        if type(val) is not int:
            raise TypeError(format_type_err_msg(self, "D_x", val, int))
        if val < 1:
            raise ValueError("D_x %d must be greater than 0." % val)
        self._D_x = val

    @property
    def D_params(self):
        return self._D_params

    @D_params.setter
    def D_params(self, val):
        if type(val) is not int:
            raise TypeError(format_type_err_msg(self, "D_params", val, int))
        if val < 1:
            raise ValueError("D_params %d must be greater than 0." % val)
        self._D_params = val

    @property
    def hidden_layers(self):
        return self._hidden_layers

    @hidden_layers.setter
    def hidden_layers(self, val):
        if type(val) is not list:
            raise TypeError(format_type_err_msg(self, "hidden_layers", val, list))
        for i, width in enumerate(val):
            if type(width) is not int:
                raise TypeError(format_type_err_msg(self, "hidden_layers[%d]" % i, val, int))
            if width < 1:
                raise ValueError("Hidden unit counts must be positive.")
        self._hidden_layers = val

    def __call__(self, x, N=100, freeze_bn=False):
        params = self.param_net(x)
        return self.density_estimator(N=N, params=params, freeze_bn=freeze_bn)

    def log_prob(self, z, x):
        params = self.param_net(x)
        return self.density_estimator.log_prob(z, params)
