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


class _WideLinearFn(torch.autograd.Function):
    """``h @ W^T + b`` of the hyper-network's LAST Linear (H -> D_params, the (M, D_params) parameter matrix of regime B)
    as ONE GEMM in the forward: the bias rides as a column of ``W`` against a ones column of ``h``.  torch's
    ``F.linear`` adds the bias in a separate epilogue pass over the M x D_params output and reduces the bias gradient in
    another one (at C4: 2.0 + 0.4 ms of a 10.3 ms training step, profiles/r02_lines/r02n_train_breakdown_final.txt).  The
    augmented width is padded to a multiple of 8."""

    @staticmethod
    def forward(ctx, h, weight, bias):
        M, H = h.shape
        Hp = (H + 1 + 7) // 8 * 8
        h_aug = h.new_zeros((M, Hp))
        h_aug[:, :H] = h
        h_aug[:, H] = 1.0
        w_aug = weight.new_zeros((weight.shape[0], Hp))
        w_aug[:, :H] = weight
        w_aug[:, H] = bias
        ctx.save_for_backward(h_aug, weight)
        ctx.H = H
        return torch.mm(h_aug, w_aug.t())

    @staticmethod
    def backward(ctx, g):
        h_aug, weight = ctx.saved_tensors
        H = ctx.H
        g = g.contiguous()
        g_h = torch.mm(g, weight) if ctx.needs_input_grad[0] else None
        # measured at C4 (profiles/r02_lines/r02zk_*): the 72-column product g^T (h | 1) picks a kernel 1.9x slower than
        # the 64-column one, so the bias gradient is a column sum here; the forward keeps the augmented GEMM (no epilogue pass)
        g_w = torch.mm(g.t(), h_aug[:, :H])
        return g_h, g_w, g.sum(dim=0)


WIDE_LINEAR_MIN_ROWS = 4096


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
        if type(val) not in (de.NormFlow, de.MoG):
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

    def _params(self, x):
        """``param_net(x)`` (reference :94,:102).  With many contexts on the GPU the last Linear runs as one GEMM per
        direction (``_WideLinearFn``); the modules and their ``state_dict`` are untouched."""
        last = self.param_net[-1]
        if (x.is_cuda and x.dim() == 2 and x.shape[0] >= WIDE_LINEAR_MIN_ROWS and x.dtype == torch.float32
                and isinstance(last, torch.nn.Linear) and last.bias is not None and len(self.param_net) >= 2):
            return _WideLinearFn.apply(self.param_net[:-1](x), last.weight, last.bias)
        return self.param_net(x)

    def __call__(self, x, N=100, freeze_bn=False):
        params = self._params(x)
        if type(self.density_estimator) is de.NormFlow:       # only the flow has BatchNorm state to freeze (:95-98)
            return self.density_estimator(N=N, params=params, freeze_bn=freeze_bn)
        return self.density_estimator(N=N, params=params)

    def log_prob(self, z, x):
        lp = self._log_prob_fused(z, x)
        if lp is not None:
            return lp
        params = self._params(x)
        return self.density_estimator.log_prob(z, params)

    # ---- hyper-network fusion: the last Linear is evaluated inside the flow kernel ------------------
    def _log_prob_fused(self, z, x):
        """log q(z | x) through ``tnf_cde_logprob`` or None when the call needs the unfused path (autograd, several
        samples per context, dropout, a chain without a compiled shape).  ``h = param_net[:-1](x)`` stays torch (a few
        small GEMMs); the (M, D_params) matrix the reference builds at conditional_density_estimator.py:102 is never
        materialised."""
        from . import _lib, config, ops
        nf = self.density_estimator
        last = self.param_net[-1]
        if (type(nf) is not de.NormFlow or not config.cde_fusion() or self.dropout or not isinstance(last, torch.nn.Linear) or len(self.param_net) < 2
                or z.dim() != 3 or z.shape[1] != 1 or x.dim() != 2 or x.shape[0] != z.shape[0]
                or z.dtype != torch.float32 or x.dtype != torch.float32 or not torch.cuda.is_available()):
            return None
        if torch.is_grad_enabled() and (z.requires_grad or x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return None
        M, _, D = z.shape
        H = last.in_features
        dev = torch.device("cuda", torch.cuda.current_device())
        pod = nf._chain_pod(_NoParams(dev, M), de._Rows(M, 0, torch.float32), None, sample=False)
        if pod is None:
            return None
        arr, keep, _ = pod
        lib = _lib.lib()
        if not lib.tnf_cde_supported(arr, len(nf.bijectors), D, H):
            return None
        home = z.device
        with torch.no_grad():
            h = ops.to_device(self.param_net[:-1](x)).contiguous()
            variant = 0 if config.cde_variant() == "tc" else 1
            key = (last.weight, last.weight._version, last.bias, last.bias._version, dev, variant)
            hit = getattr(self, "_cde_pack", None)
            if hit is None or any(a is not b if isinstance(a, torch.Tensor) else a != b for a, b in zip(hit[0], key)):
                w = ops.to_device(last.weight.detach()).contiguous()
                b = ops.to_device(last.bias.detach()).contiguous()
                packed = torch.empty(lib.tnf_cde_packed_bytes(nf.D_params, H, variant), dtype=torch.uint8, device=dev)
                _lib.check(lib.tnf_cde_pack(arr, len(nf.bijectors), D, w.data_ptr(), b.data_ptr(), H, packed.data_ptr(),
                                            variant, ops._stream()), "tnf_cde_pack")
                self._cde_pack = hit = (key, packed)
            zd = ops.to_device(z).contiguous()
            lp = torch.empty((M, 1), dtype=torch.float32, device=dev)
            _lib.check(lib.tnf_cde_logprob(arr, len(nf.bijectors), D, h.data_ptr(), H, hit[1].data_ptr(), zd.data_ptr(), M,
                                           lp.data_ptr(), variant, ops._stream()), "tnf_cde_logprob")
        return de._to(lp, home)


class _NoParams(object):
    """Stand-in for the parameter matrix in ``NormFlow._chain_pod`` (device and row count are all it reads there)."""

    def __init__(self, device, M):
        self.device, self.shape = device, (max(2, M), 0)
