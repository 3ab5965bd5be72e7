"""Bijectors of the torch_nf API, executed by sm_100a CUDA kernels.

Host-side mirror of reference ``torch_nf/bijectors.py``: same class names,
constructor arguments, validation / clamping behaviour, ``name`` / ``D``
attributes, ``count_num_params`` and the flat ``(M, D_params)`` parameter
packing, so objects are drop-in replacements behind ``NormFlow``.  All
arithmetic is done by the C-ABI kernels (``include/tnf.h``); tensors handed in
on the CPU are staged to the GPU and the results returned on the caller's
device.  There is no CPU compute path.
"""
import numpy as np
import torch

from . import config, ops
from .error_formatters import format_type_err_msg
from .ops import TNF_FORWARD, TNF_INVERSE


def _stage(z, params=None):
    """Stage inputs on the GPU. Returns (z_dev, params_dev, home_device)."""
    if not isinstance(z, torch.Tensor):
        raise TypeError("z must be a torch.Tensor, got %s" % type(z).__name__)
    home = z.device
    if z.dtype not in (torch.float32, torch.float64):
        z = z.float()
    zd = ops.to_device(z)
    pd = None
    if params is not None:
        pd = ops.to_device(params, zd.dtype)
    return zd, pd, home


def _home(t, home):
    return ops.to_like(t, home)


class Bijector(object):
    """Base class (reference bijectors.py:7-71)."""

    def __init__(self, D):
        super().__init__()
        self.D = D

    @property
    def D(self):
        return self._D

    @D.setter
    def D(self, val):
        if type(val) is not int:
            raise TypeError(format_type_err_msg(self, "D", val, int))
        if val < 1:
            raise ValueError("Bijector dimensionality must be positive.")
        self._D = val

    def __call__(self, z, params):
        return self.forward_and_log_det(z, params)

    def forward_and_log_det(self, z, params):
        raise NotImplementedError()

    def inverse_and_log_det(self, z, params):
        raise NotImplementedError()

    def count_num_params(self):
        return 0


# ------------------------------------------------------------------ RealNVP
def _tc_differentiable(z, params, D, U, L):
    """Shared weights, a compiled shape and the tensor-core backward enabled (config.tc_backward_enabled: the bf16 mode, or
    set_training_backward("bf16")): the layer AND its backward run on tensor cores (tnf_coupling_tc /
    tnf_coupling_tc_bwd); otherwise the exact CUDA-core kernels."""
    return (config.tc_backward_enabled() and params.shape[0] == 1 and z.dtype == torch.float32
            and params.dtype == torch.float32 and z.shape[0] * z.shape[1] >= config.tc_min_rows()
            and ops.tc_bwd_supported(D, U, L) and ops.tc_supported(D, U, L, config.tc_precision()))


class _CouplingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, params, D, U, L, upper, direction):
        ctx.set_materialize_grads(False)
        ctx.tc = (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) and _tc_differentiable(z, params, D, U, L)
        if ctx.tc:
            zc = z.contiguous()
            mode = config.tc_precision()
            z_out, ld = ops.coupling_tc(zc, ops.tc_pack(params[0], D, U, L, upper, precision=mode), D, U, L, upper,
                                        direction, precision=mode)
            ld = ld.view(z.shape[0], z.shape[1])
        else:
            z_out, ld = ops.coupling(z, params, D, U, L, upper, direction)
        ctx.save_for_backward(z, params)
        ctx.cfg = (D, U, L, upper, direction)
        return z_out, ld

    @staticmethod
    def backward(ctx, g_z, g_ld):
        z, params = ctx.saved_tensors
        D, U, L, upper, direction = ctx.cfg
        g_params = torch.zeros(params.shape, dtype=params.dtype, device=params.device)
        if ctx.tc:
            n = ops.coupling_num_params(D, U, L, upper)
            g_in = ops.coupling_tc_bwd(z.contiguous(), ops.tc_bwd_pack(params[0], D, U, L, upper), g_z, g_ld,
                                       g_params[0, :n], D, U, L, upper, direction)
        else:
            g_in = ops.coupling_bwd(z, params, g_z, g_ld, g_params, D, U, L, upper, direction)
        return g_in, g_params, None, None, None, None, None


class RealNVP(Bijector):
    """Affine coupling layer (reference bijectors.py:74-262).

    ``transform_upper=True`` transforms columns ``D//2:`` conditioned on
    columns ``:D//2``.  Parameter packing per conditioner layer, from the front
    of ``params``: ``W_t (K*J)``, ``W_s (K*J)``, ``b_t (J)``, ``b_s (J)`` with
    ``W`` viewed ``(K, J)`` (``x @ W``); layers ``D_in->U``, ``U->U`` x (L-1),
    ``U->D_out``; tanh on all but the last (bijectors.py:168-171,224-242).
    """

    def __init__(self, D, num_layers, num_units, transform_upper=True):
        super().__init__(D)
        self.name = "RealNVP"
        self.num_layers = num_layers
        self.num_units = num_units
        self.transform_upper = transform_upper

    @property
    def num_layers(self):
        return self._num_layers

    @num_layers.setter
    def num_layers(self, val):
        if type(val) is not int:
            raise TypeError(format_type_err_msg(self, "num_layers", val, int))
        if val < 1:
            raise ValueError("RealNVP.num_layers must be positive.")
        if val > 5:
            print("Warning: RealNVP.num_layers set to maximum of 5 (received %d)." % val)
            val = 5
        self._num_layers = val

    @property
    def num_units(self):
        return self._num_units

    @num_units.setter
    def num_units(self, val):
        if type(val) is not int:
            raise TypeError(format_type_err_msg(self, "num_units", val, int))
        if val < 15:
            print("Warning: num_units set to minimum of 15 (received %d)." % val)
            val = 15
        elif val > 1000:
            print("Warning: num_units set to maximum of 1,000 (received %d)." % val)
            val = 1000
        self._num_units = val

    @property
    def transform_upper(self):
        return self._transform_upper

    @transform_upper.setter
    def transform_upper(self, val):
        if type(val) is not bool:
            raise TypeError(format_type_err_msg(self, "transform_upper", val, bool))
        self._transform_upper = val

    def _dims(self):
        h = self.D // 2
        if self.transform_upper:
            return h, self.D - h
        return self.D - h, h

    def count_num_params(self):
        """bijectors.py:244-262."""
        d_in, d_out = self._dims()
        U = self.num_units
        return 2 * (d_in * U + d_out * U + d_out + U + (self.num_layers - 1) * (U + 1) * U)

    def _run(self, z, params, direction):
        zd, pd, home = _stage(z, params)
        if zd.dim() != 3 or zd.shape[2] != self.D:
            raise ValueError("RealNVP expects z of shape (M, N, %d), got %s" % (self.D, tuple(z.shape)))
        if pd.dim() != 2 or pd.shape[1] < self.count_num_params():
            raise ValueError("RealNVP needs %d parameters per row, got %s" % (self.count_num_params(), tuple(params.shape)))
        z_out, ld = _CouplingFn.apply(zd, pd, self.D, self.num_units, self.num_layers, self.transform_upper, direction)
        return _home(z_out, home), _home(ld, home)

    def forward_and_log_det(self, z, params):
        """z2' = t + z2*exp(s), log_det = sum(s)  (bijectors.py:145-181)."""
        return self._run(z, params, TNF_FORWARD)

    def inverse_and_log_det(self, z, params):
        """z2' = (z2 - t)/exp(s), returns +sum(s)  (bijectors.py:183-206)."""
        return self._run(z, params, TNF_INVERSE)


# ------------------------------------------------------------------ Affine
class _AffineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, params, D, direction):
        ctx.set_materialize_grads(False)
        z_out, ld = ops.affine(z, params, D, direction)
        ctx.save_for_backward(z, params)
        ctx.cfg = (D, direction)
        return z_out, ld

    @staticmethod
    def backward(ctx, g_z, g_ld):
        z, params = ctx.saved_tensors
        D, direction = ctx.cfg
        g_params = torch.zeros(params.shape, dtype=params.dtype, device=params.device)
        g_in = ops.affine_bwd(z, params, g_z, g_ld, g_params, D, direction)
        return g_in, g_params, None, None


class Affine(Bijector):
    """Per-dimension scale and shift, params = [alpha(D), shift(D)]
    (reference bijectors.py:265-318). ``log_det`` has shape ``(M, 1)``."""

    def __init__(self, D):
        super().__init__(D)
        self.name = "Affine"

    def _run(self, z, params, direction):
        zd, pd, home = _stage(z, params)
        if zd.dim() != 3 or zd.shape[2] != self.D:
            raise ValueError("Affine expects z of shape (M, N, %d), got %s" % (self.D, tuple(z.shape)))
        z_out, ld = _AffineFn.apply(zd, pd, self.D, direction)
        return _home(z_out, home), _home(ld, home)

    def forward_and_log_det(self, z, params):
        return self._run(z, params, TNF_FORWARD)

    def inverse_and_log_det(self, z, params):
        return self._run(z, params, TNF_INVERSE)

    def count_num_params(self):
        return 2 * self.D


# ------------------------------------------------------------------ BatchNorm
def _reduce_stats(sums):
    """Data-parallel hook: all-reduce [sum | sumsq | rows] across the sample shards."""
    from . import dist
    return dist.allreduce_stats(sums)


class _BatchNormFn(torch.autograd.Function):
    """Batch-statistics normalisation y = (z - mean)/alpha with
    log_det = -sum log alpha; differentiable through the statistics."""

    @staticmethod
    def forward(ctx, z, D, eps):
        ctx.set_materialize_grads(False)
        sums = _reduce_stats(ops.colstats(z, D))
        mean, alpha, ld = ops.bn_finalize(sums, D, eps, z.dtype)
        y = ops.bn_apply(z, mean, alpha, D, TNF_FORWARD)
        ctx.save_for_backward(y, alpha, sums)
        ctx.D = D
        ctx.mark_non_differentiable(mean, alpha)
        return y, ld, mean, alpha

    @staticmethod
    def backward(ctx, g_y, g_ld, _gm, _ga):
        y, alpha, sums = ctx.saved_tensors
        D = ctx.D
        if g_y is None:
            g_y = torch.zeros_like(y)
        gs = _reduce_stats(ops.bn_bwd_sums(g_y, y, D))
        g_z = ops.bn_bwd_apply(g_y, y, alpha, gs, g_ld, sums[2 * D:], D)
        return g_z, None, None


class _BnApplyFn(torch.autograd.Function):
    """Normalisation with stored (constant) statistics, either direction."""

    @staticmethod
    def forward(ctx, z, mean, alpha, D, direction):
        ctx.save_for_backward(alpha)
        ctx.cfg = (D, direction)
        return ops.bn_apply(z, mean, alpha, D, direction)

    @staticmethod
    def backward(ctx, g):
        (alpha,) = ctx.saved_tensors
        D, direction = ctx.cfg
        # d/dz (z-mean)/alpha = 1/alpha ; d/dz (z*alpha+mean) = alpha: the same kernel with mean = 0
        g_in = ops.bn_apply(g.contiguous(), torch.zeros_like(alpha), alpha, D, direction)
        return g_in, None, None, None, None


class BatchNorm(Bijector):
    """Batch normalisation with log-det (reference bijectors.py:321-426).

    ``__call__(z, use_last=False)`` takes no ``params``.  With
    ``use_last=False`` the statistics of the whole flattened ``(M*N, D)``
    batch are used and remembered; ``use_last=True`` and
    ``inverse_and_log_det`` use the remembered ones (initially mean 0,
    alpha 1).  ``alpha = sqrt(biased var + eps)``.  When a data-parallel group
    is active (``torch_nf_b200.dist``) the statistics are all-reduced so every
    shard normalises with the global batch statistics.

    Difference from the reference: the remembered statistics are detached
    constants (the reference keeps their autograd graph alive).
    """

    def __init__(self, D, momentum=0.1, eps=1e-5):
        super().__init__(D)
        self.name = "BatchNorm"
        self.momentum = momentum
        self.eps = eps
        self._last_mean = torch.zeros(D)
        self._last_alpha = torch.ones(D)
        self._last_ld = torch.zeros(())   # -sum(log alpha) of the remembered statistics
        self._home_dev = torch.device("cpu")

    @property
    def momentum(self):
        return self._momentum

    @momentum.setter
    def momentum(self, val):
        if type(val) is not float:
            raise TypeError(format_type_err_msg(self, "momentum", val, float))
        if val < 0.0:
            raise ValueError("BatchNorm.momentum cannot be negative.")
        if val > 1.0:
            print("Warning: BathNorm.momentum  set to maximum of 1.0 (received %.2E)." % val)
            val = 1.0
        self._momentum = val

    @property
    def eps(self):
        return self._eps

    @eps.setter
    def eps(self, val):
        if type(val) is not float:
            raise TypeError(format_type_err_msg(self, "eps", val, float))
        if val < 0.0:
            raise ValueError("BatchNorm.eps cannot be negative.")
        self._eps = val

    def get_last_mean(self):
        return self._last_mean.to(self._home_dev)

    def get_last_alpha(self):
        return self._last_alpha.to(self._home_dev)

    def _state_on(self, device, dtype):
        m, a = self._last_mean, self._last_alpha
        if m.device != device or m.dtype != dtype:
            m, a = m.to(device=device, dtype=dtype), a.to(device=device, dtype=dtype)
        return m.contiguous(), a.contiguous()

    def __call__(self, z, use_last=False):
        return self.forward_and_log_det(z, use_last=use_last)

    def forward_and_log_det(self, z, use_last=False):
        zd, _, home = _stage(z)
        if zd.shape[-1] != self.D:
            raise ValueError("BatchNorm expects last dimension %d, got %s" % (self.D, tuple(z.shape)))
        zd = zd.contiguous()
        if use_last:
            mean, alpha = self._state_on(zd.device, zd.dtype)
            y = _BnApplyFn.apply(zd, mean, alpha, self.D, TNF_FORWARD)
            ld = self._last_ld.to(device=zd.device, dtype=zd.dtype)
        else:
            y, ld, mean, alpha = _BatchNormFn.apply(zd, self.D, self.eps)
            self._set_state(mean.detach(), alpha.detach(), ld.detach(), home)
        return _home(y, home), _home(ld, home)

    def inverse_and_log_det(self, z):
        zd, _, home = _stage(z)
        zd = zd.contiguous()
        mean, alpha = self._state_on(zd.device, zd.dtype)
        x = _BnApplyFn.apply(zd, mean, alpha, self.D, TNF_INVERSE)
        return _home(x, home), self._last_ld.to(device=home, dtype=zd.dtype)

    def _set_state(self, mean, alpha, ld, home=None):
        self._last_mean, self._last_alpha, self._last_ld = mean, alpha, ld
        if home is not None:
            self._home_dev = home


# ------------------------------------------------------------------ ToInterval
class _ToIntervalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, consts, D, direction):
        ctx.set_materialize_grads(False)
        z_out, ld = ops.tointerval(z, consts, D, direction)
        ctx.save_for_backward(z, consts)
        ctx.cfg = (D, direction)
        return z_out, ld

    @staticmethod
    def backward(ctx, g_z, g_ld):
        z, consts = ctx.saved_tensors
        D, direction = ctx.cfg
        return ops.tointerval_bwd(z, consts, g_z, g_ld, D, direction), None, None, None


class ToInterval(Bijector):
    """Maps each dimension into [lb, ub] / [lb, inf) / (-inf, ub] / R
    (reference bijectors.py:429-557): tanh for two-sided bounds, softplus for
    one-sided ones, identity otherwise.  ``lb`` / ``ub`` must be lists or
    numpy arrays of equal length with ``lb <= ub``."""

    def __init__(self, D, lb, ub):
        super().__init__(D)
        self.name = "ToInterval"
        self.lb = lb
        self.ub = ub
        self._eps = 1e-12
        if self.lb.shape[0] != self.ub.shape[0]:
            raise ValueError("Lower and upper bounds must be same length.")
        for lo, hi in zip(self.lb, self.ub):
            if lo > hi:
                raise ValueError("Lower bound %.2E > upper bound %.2E." % (lo, hi))
        c = np.zeros((6, self.D), dtype=np.float64)   # tanh_flg, sp_flg, tanh_m, tanh_c, sp_m, sp_c
        c[2] = 1.0
        c[4] = 1.0
        for i in range(self.D):
            lo, hi = self.lb[i], self.ub[i]
            has_lo, has_hi = not np.isneginf(lo), not np.isposinf(hi)
            if has_lo and has_hi:
                c[0, i], c[2, i], c[3, i] = 1.0, (hi - lo) / 2.0, (hi + lo) / 2.0
            elif has_lo:
                c[1, i], c[4, i], c[5, i] = 1.0, 1.0, lo
            elif has_hi:
                c[1, i], c[4, i], c[5, i] = 1.0, -1.0, hi
        c32 = torch.tensor(c).float()
        # same attribute names / shapes as the reference (bijectors.py:475-480)
        self.tanh_flg, self.softplus_flg = c32[0][None, None, :], c32[1][None, None, :]
        self.tanh_m, self.tanh_c = c32[2][None, None, :], c32[3][None, None, :]
        self.softplus_m, self.softplus_c = c32[4][None, None, :], c32[5][None, None, :]
        # 7th row: log(tanh_m) in float32, evaluated once on the host exactly as the reference does (:515)
        self._consts_cpu = torch.cat([c32, torch.log(c32[2])[None, :]], dim=0).contiguous()
        self._consts_dev = {}

    @property
    def lb(self):
        return self._lb

    @lb.setter
    def lb(self, val):
        if type(val) not in [list, np.ndarray]:
            raise TypeError(format_type_err_msg(self, "lb", val, np.ndarray))
        self._lb = np.array(val) if type(val) is list else val

    @property
    def ub(self):
        return self._ub

    @ub.setter
    def ub(self, val):
        if type(val) not in [list, np.ndarray]:
            raise TypeError(format_type_err_msg(self, "ub", val, np.ndarray))
        self._ub = np.array(val) if type(val) is list else val

    def _consts(self, device):
        c = self._consts_dev.get(device)
        if c is None:
            c = self._consts_cpu.to(device)
            self._consts_dev[device] = c
        return c

    def __call__(self, z):
        return self.forward_and_log_det(z)

    def _run(self, z, direction):
        zd, _, home = _stage(z)
        if zd.dim() != 3 or zd.shape[2] != self.D:
            raise ValueError("ToInterval expects z of shape (M, N, %d), got %s" % (self.D, tuple(z.shape)))
        z_out, ld = _ToIntervalFn.apply(zd.contiguous(), self._consts(zd.device), self.D, direction)
        return _home(z_out, home), _home(ld, home)

    def forward_and_log_det(self, z):
        return self._run(z, TNF_FORWARD)

    def inverse_and_log_det(self, z):
        return self._run(z, TNF_INVERSE)


def torch_atanh(x):
    """atanh with the reference's 1e-12 guards (bijectors.py:555-557); kept for
    API parity -- the ToInterval kernel evaluates the same expression."""
    _eps = 1e-12
    return 0.5 * (torch.log(1 + x + _eps) - torch.log(1 - x + _eps))


# ------------------------------------------------------------------ ToSimplex
class _ToSimplexFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, D_attr):
        ctx.set_materialize_grads(False)
        z_out, ld = ops.tosimplex(z, D_attr)
        ctx.save_for_backward(z)
        ctx.D_attr = D_attr
        return z_out, ld

    @staticmethod
    def backward(ctx, g_z, g_ld):
        (z,) = ctx.saved_tensors
        return ops.tosimplex_bwd(z, g_z, g_ld, ctx.D_attr), None


class ToSimplex(Bijector):
    """(M, N, K) -> (M, N, K+1) on the simplex (reference bijectors.py:560-594).
    Sample direction only: the reference defines no inverse."""

    def __init__(self, D):
        super().__init__(D)
        self.name = "ToSimplex"

    def __call__(self, z):
        return self.forward_and_log_det(z)

    def forward_and_log_det(self, z):
        zd, _, home = _stage(z)
        if zd.dim() != 3:
            raise ValueError("ToSimplex expects z of shape (M, N, K), got %s" % (tuple(z.shape),))
        z_out, ld = _ToSimplexFn.apply(zd.contiguous(), self.D)
        return _home(z_out, home), _home(ld, home)

    def count_num_params(self):
        return 0


# ------------------------------------------------------------------ MAF
class MAF(Bijector):
    """Masked autoregressive flow (reference bijectors.py:597-806).

    Clamps, the random masks drawn from numpy's global stream, the parameter
    count and the packing ``[W_mu, W_alpha]`` per layer (no biases) follow the
    reference, so ``NormFlow(arch_type='AR')`` constructs identically.  Both
    directions run on the exact CUDA-core coupling kernel in its MAF mode
    (every column conditions and is transformed; weights multiplied by the
    masks on load); the backward exists for the inverse (log_prob) direction.
    """

    def __init__(self, D, num_layers, num_units, fwd_fac=True):
        super().__init__(D)
        self.name = "MAF"
        self.num_layers = num_layers
        self.num_units = num_units
        self.fwd_fac = fwd_fac
        self._mask_dev = {}
        self._get_masks()

    @property
    def num_layers(self):
        return self._num_layers

    @num_layers.setter
    def num_layers(self, val):
        if type(val) is not int:
            raise TypeError(format_type_err_msg(self, "num_layers", val, int))
        if val < 1:
            raise ValueError("MAF.num_layers must be positive.")
        if val > 5:
            print("Warning: MAF.num_layers set to maximum of 5 (received %d)." % val)
            val = 5
        self._num_layers = val

    @property
    def num_units(self):
        return self._num_units

    @num_units.setter
    def num_units(self, val):
        if type(val) is not int:
            raise TypeError(format_type_err_msg(self, "num_units", val, int))
        if val < 5:
            print("Warning: num_units set to minimum of 15 (received %d)." % val)
            val = 5
        elif val > 1000:
            print("Warning: num_units set to maximum of 1,000 (received %d)." % val)
            val = 1000
        self._num_units = val

    @property
    def fwd_fac(self):
        return self._fwd_fac

    @fwd_fac.setter
    def fwd_fac(self, val):
        if type(val) is not bool:
            raise TypeError(format_type_err_msg(self, "fwd_fac", val, bool))
        self._fwd_fac = val

    def _get_masks(self):
        """Degrees ``ms`` and binary masks ``Ms`` (bijectors.py:663-696): hidden
        degrees are drawn with ``np.random.randint(1, D, K)``; a hidden unit sees
        inputs of degree <= its own, an output only hidden units of degree < its own."""
        D, K = self.D, self.num_units
        order = np.arange(1, D + 1) if self.fwd_fac else np.arange(D, -1, -1)
        self.ms, self.Ms = [], []
        prev = order[:D]
        for _ in range(self.num_layers):
            deg = np.random.randint(1, D, (K,))
            mask = (prev[:, None] <= deg[None, :]).astype(np.float64)
            self.Ms.append(torch.tensor(mask[None, :, :]).float())
            self.ms.append(deg)
            prev = deg
        mask = (prev[:, None] < order[None, :D]).astype(np.float64)
        self.ms.append(order)
        self.Ms.append(torch.tensor(mask[None, :, :]).float())
        return None

    def count_num_params(self):
        """bijectors.py:804-806."""
        return 2 * (2 * self.D * self.num_units + (self.num_layers - 1) * (self.num_units ** 2))

    def _mask_flat(self, device):
        """The masks in the parameter-row layout: per layer [mask, mask] (W_mu and W_alpha share it)."""
        m = self._mask_dev.get(device)
        if m is None:
            parts = []
            for Mi in self.Ms:
                flat = Mi.reshape(-1).float()
                parts += [flat, flat]
            m = torch.cat(parts).contiguous().to(device)
            self._mask_dev[device] = m
        return m

    def _run(self, z, params, direction):
        zd, pd, home = _stage(z, params)
        if zd.dim() != 3 or zd.shape[2] != self.D:
            raise ValueError("MAF expects z of shape (M, N, %d), got %s" % (self.D, tuple(z.shape)))
        if pd.dim() != 2 or pd.shape[1] < self.count_num_params():
            raise ValueError("MAF needs %d parameters per row, got %s" % (self.count_num_params(), tuple(params.shape)))
        z_out, ld = _MafFn.apply(zd, pd, self._mask_flat(zd.device), self.D, self.num_units, self.num_layers, direction)
        return _home(z_out, home), _home(ld, home)

    def forward_and_log_det(self, z, params):
        """D-1 passes z <- u*exp(alpha(z)) + mu(z) (bijectors.py:742-756)."""
        return self._run(z, params, TNF_FORWARD)

    def inverse_and_log_det(self, z, params):
        """z' = (z - mu(z))/exp(alpha(z)) in one pass (bijectors.py:758-764)."""
        return self._run(z, params, TNF_INVERSE)


class _MafFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, params, mask, D, U, L, direction):
        ctx.set_materialize_grads(False)
        z_out, ld = ops.maf(z, params, mask, D, U, L, direction)
        ctx.save_for_backward(z, params, mask)
        ctx.cfg = (D, U, L, direction)
        return z_out, ld

    @staticmethod
    def backward(ctx, g_z, g_ld):
        z, params, mask = ctx.saved_tensors
        D, U, L, direction = ctx.cfg
        if direction != TNF_INVERSE:
            raise NotImplementedError("MAF backward is built for the inverse (log_prob) direction only")
        g_params = torch.zeros(params.shape, dtype=params.dtype, device=params.device)
        g_in = ops.maf_bwd(z, params, mask, g_z, g_ld, g_params, D, U, L, direction)
        return g_in, g_params, None, None, None, None, None
