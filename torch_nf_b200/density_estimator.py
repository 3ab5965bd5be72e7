"""Density estimators of the torch_nf API on the B200 hot path.

Host-side mirror of reference ``torch_nf/density_estimator.py``
(``DensityEstimator`` :11-55, ``NormFlow`` :240-421): same constructor,
validation, bijector order, parameter slicing and return dtypes
(``forward`` -> ``z`` float32 + ``log_q_z`` float64, ``log_prob`` -> float32).
The chain runs as CUDA kernels; without autograd it runs as a fused plan in
which every bijector accumulates its log-det in place, BatchNorm / Affine are
folded into the neighbouring coupling kernel where the tensor-core path is
active, and nothing but ``z`` and the accumulators touches HBM between layers.
"""
import numpy as np
import torch

from . import config, ops
from .bijectors import Affine, BatchNorm, Bijector, MAF, RealNVP, ToInterval, ToSimplex
from .error_formatters import format_type_err_msg
from .ops import TNF_FORWARD, TNF_INVERSE, TNF_LD_ADD


class DensityEstimator(object):
    """Abstract base (reference density_estimator.py:11-55)."""

    def __init__(self, D, conditioner=False):
        super().__init__()
        self.D = D
        self.conditioner = conditioner

    @property
    def D(self):
        return self._D

    @D.setter
    def D(self, val):
        if type(val) is not int:
            raise TypeError(format_type_err_msg(self, "D", val, int))
        if val < 2:
            raise ValueError("DensityEstimator D %d must be greater than 1." % val)
        self._D = val

    @property
    def conditioner(self):
        return self._conditioner

    @conditioner.setter
    def conditioner(self, val):
        if type(val) is not bool:
            raise TypeError(format_type_err_msg(self, "conditioner", val, bool))
        self._conditioner = val

    def __call__(self, N=100, params=None):
        if not self.conditioner:
            return self.forward(self.params, N)
        return self.forward(params, N)

    def forward(self, params, N=100, freeze_bn=False):
        raise NotImplementedError()

    def log_prob(self, z, params=None):
        raise NotImplementedError()

    def count_num_params(self):
        raise NotImplementedError()

    def _param_init(self):
        raise NotImplementedError()


MOG_EPS = 1e-12


class MoG(DensityEstimator):
    """Mixture of K Gaussians (reference density_estimator.py:57-237): the other density estimator a
    ``ConditionalDensityEstimator`` accepts.  No bijector chain - it is adjacent to the hot path and is
    evaluated with torch ops: the parameter map (``_get_MoG_params``, a few values per context) on the device of
    ``params``; sampling and the densities, which scale with ``M N K D^2``, on the CUDA device.  Parameter row:
    ``[alpha logits (K) | mu (K D) | upper-triangular factor U (K D(D+1)/2)]`` with ``Sigma^-1 = U^T U`` and
    ``diag(U) = exp(.)`` (:98-133); with bounds ``mu = m tanh(mu) + c`` and ``diag(U) /= sqrt(m)``.

    Difference from the reference: ``forward`` draws its samples on the device (component index by inverse-CDF,
    ``mu_k + chol(Sigma_k + 1e-3 I) eps``) instead of one scipy call per sample (:146-155) - the same distribution,
    another random stream; ``log_q_z`` is the reference's float64 mixture density of the drawn points (:160, :217-235)."""

    def __init__(self, D, conditioner=False, K=1, lb=None, ub=None):
        super().__init__(D, conditioner)
        self.K = K
        self.alpha_softmax = torch.nn.Softmax(dim=1)
        self.count_num_params()
        if not self.conditioner:
            self._param_init()
        self.lb = lb
        self.ub = ub

    @property
    def K(self):
        return self._K

    @K.setter
    def K(self, val):
        if type(val) is not int:
            raise TypeError(format_type_err_msg(self, "K", val, int))
        if val < 1:
            raise ValueError("MoG K %d must be greater than 0." % val)
        self._K = val

    def count_num_params(self):
        self.D_params = self.K * (1 + self.D + self.D * (self.D + 1) // 2)

    def _param_init(self):
        self.params = torch.nn.init.xavier_normal_(torch.zeros(1, self.D_params, requires_grad=True))

    def _bounds(self, like):
        if self.lb is None or self.ub is None:
            return None, None
        lb, ub = np.asarray(self.lb, dtype=np.float64), np.asarray(self.ub, dtype=np.float64)
        m = torch.tensor((ub - lb) / 2.0).float().to(like.device)
        c = torch.tensor((ub + lb) / 2.0).float().to(like.device)
        return m, c

    def _get_MoG_params(self, params, numpy=False):
        """alpha (M, K), mu (M, K, D), Sigma_inv (M, K, D, D), Sigma_det (M, K) (:89-140)."""
        M, K, D = params.shape[0], self.K, self.D
        T = D * (D + 1) // 2
        m, c = self._bounds(params)
        alpha = self.alpha_softmax(params[:, :K])
        mu = params[:, K:K + K * D].reshape(M, K, D)
        if m is not None:
            mu = m * torch.tanh(mu) + c
        tri = params[:, K + K * D:K + K * D + K * T].reshape(M, K, T)
        iu = torch.triu_indices(D, D, device=params.device)
        is_diag = iu[0] == iu[1]
        log_diag = tri[:, :, is_diag]                                  # (M, K, D): row-major triu order visits (d, d) in order
        diag = torch.exp(log_diag)
        if m is not None:
            diag = diag / torch.sqrt(m)
        vals = tri.clone()
        vals[:, :, is_diag] = diag
        U = torch.zeros((M, K, D, D), dtype=params.dtype, device=params.device)
        U[:, :, iu[0], iu[1]] = vals
        Sigma_inv = torch.matmul(U.transpose(3, 2), U)
        Sigma_det = torch.prod(torch.exp(-2.0 * log_diag) * (m if m is not None else 1.0), dim=2)
        if numpy:
            alpha = alpha.detach().cpu().numpy()
            alpha = alpha / np.sum(alpha, axis=1)[:, None]
            mu = mu.detach().cpu().numpy()
            Sigma_inv = Sigma_inv.detach().cpu().numpy()
        return alpha, mu, Sigma_inv, Sigma_det

    def _mixture_log_density64(self, z, alpha, mu, Sigma_inv):
        """log(sum_k alpha_k N(z; mu_k, Sigma_k) + EPS) in float64 (the reference's ``log_prob_np``, :217-235)."""
        z, alpha, mu, P = z.double(), alpha.double(), mu.double(), Sigma_inv.double()
        alpha = alpha / alpha.sum(dim=1, keepdim=True)
        d = z[:, :, None, :] - mu[:, None, :, :]                                        # (M, N, K, D)
        quad = torch.einsum("mnki,mkij,mnkj->mnk", d, P, d)
        log_norm = 0.5 * torch.logdet(P) - 0.5 * self.D * float(np.log(2.0 * np.pi))     # (M, K)
        pdf = torch.exp(log_norm[:, None, :] - 0.5 * quad)
        return torch.log((alpha[:, None, :] * pdf).sum(dim=2) + MOG_EPS)

    def forward(self, params, N=100):
        """Samples ``z (M, N, D)`` and ``log_q_z (M, N)``, float32, on the device of ``params`` (:142-165)."""
        home = params.device
        pd = ops.to_device(params.detach(), torch.float32)
        M, K, D = pd.shape[0], self.K, self.D
        alpha, mu, Sigma_inv, _ = self._get_MoG_params(pd)
        alpha = alpha / alpha.sum(dim=1, keepdim=True)
        Sigma = torch.inverse(Sigma_inv.double()) + 0.001 * torch.eye(D, dtype=torch.float64, device=pd.device)
        L = torch.linalg.cholesky(Sigma)                                                # (M, K, D, D)
        g = torch.Generator(device=pd.device)
        g.manual_seed(int(np.random.randint(0, 2 ** 31 - 1)))      # numpy's global stream seeds the device stream
        u = torch.rand((M, N), generator=g, device=pd.device, dtype=torch.float64)
        comp = torch.searchsorted(torch.cumsum(alpha.double(), dim=1).contiguous(), u).clamp_(max=K - 1)   # (M, N)
        eps = torch.randn((M, N, D), generator=g, device=pd.device, dtype=torch.float64)
        rows = torch.arange(M, device=pd.device)[:, None].expand(M, N)
        z = mu.double()[rows, comp] + torch.einsum("mnij,mnj->mni", L[rows, comp], eps)
        log_q_z = self._mixture_log_density64(z, alpha, mu, Sigma_inv)
        return _to(z.float(), home), _to(log_q_z.float(), home)

    def log_prob(self, z, params=None):
        """log q(z) with the reference's two formulas (K = 1: exact Gaussian log-density; K > 1: log of the summed
        component densities + EPS, :167-215), float32, on the device of ``z``."""
        if params is None:
            params = self.params
        home = z.device
        zd = ops.to_device(z, torch.float32)
        pd = ops.to_device(params, torch.float32)
        alpha, mu, Sigma_inv, Sigma_det = self._get_MoG_params(pd)
        d = zd[:, :, None, :] - mu[:, None, :, :]                                       # (M, N, K, D)
        quad = torch.einsum("mnki,mkij,mnkj->mnk", d, Sigma_inv, d)
        if self.K == 1:
            lp = quad[:, :, 0] + torch.log(Sigma_det + MOG_EPS) + self.D * float(np.log(2.0 * np.pi))
            lp = -0.5 * lp
        else:
            denom = torch.sqrt(((2.0 * np.pi) ** self.D) * Sigma_det + MOG_EPS)[:, None, :]
            prob = torch.sum(alpha[:, None, :] * (torch.exp(-0.5 * quad) / denom), dim=2)
            lp = torch.log(prob + MOG_EPS)
        return _to(lp, home)

    def log_prob_np(self, z, params):
        """numpy in / out variant (:217-235)."""
        pd = ops.to_device(params.detach(), torch.float32)
        alpha, mu, Sigma_inv, _ = self._get_MoG_params(pd)
        zt = ops.to_device(torch.as_tensor(np.asarray(z)), torch.float64)
        return self._mixture_log_density64(zt, alpha, mu, Sigma_inv).cpu().numpy()


class NormFlow(DensityEstimator):
    """Normalizing flow (reference density_estimator.py:240-421).

    ``arch_type='coupling'`` stacks, per stage,
    ``RealNVP(upper) -> BatchNorm -> RealNVP(lower) -> BatchNorm -> Affine``
    (:260-270); ``'AR'`` is ``MAF -> BatchNorm -> Affine``; ``'affine'`` a single
    ``Affine``; an optional ``support_layer`` bijector instance is appended.
    ``params`` is the flat ``(M, D_params)`` matrix in chain order.

    Extensions that do not change reference behaviour: ``forward(...,
    omega=)`` injects the base noise (parity runs); tensors may live on the GPU,
    in which case results stay there.
    """

    def __init__(self, D, conditioner=False, arch_type="AR", num_stages=1, num_layers=2, num_units=15,
                 support_layer=None):
        super().__init__(D, conditioner)
        self.arch_type = arch_type
        self.num_stages = num_stages
        self.num_layers = num_layers
        self.num_units = num_units
        self.support_layer = support_layer

        self.bijectors = []
        if arch_type == "coupling":
            for _ in range(num_stages):
                self.bijectors.append(RealNVP(D, num_layers, num_units, transform_upper=True))
                self.bijectors.append(BatchNorm(D))
                self.bijectors.append(RealNVP(D, num_layers, num_units, transform_upper=False))
                self.bijectors.append(BatchNorm(D))
                self.bijectors.append(Affine(D))
        elif arch_type == "AR":
            self.bijectors.append(MAF(D, self.num_layers, self.num_units, fwd_fac=True))
            self.bijectors.append(BatchNorm(D))
            self.bijectors.append(Affine(D))
        elif arch_type == "affine":
            self.bijectors.append(Affine(D))

        if support_layer is not None:
            if issubclass(type(support_layer), Bijector):
                self.bijectors.append(support_layer)
            else:
                raise TypeError("Support layer not Bijector.")

        self.count_num_params()
        self._tc_cache = {}
        if not self.conditioner:
            self._param_init()

    # ---- validated attributes (density_estimator.py:289-350)
    @property
    def arch_type(self):
        return self._arch_type

    @arch_type.setter
    def arch_type(self, val):
        if type(val) is not str:
            raise TypeError(format_type_err_msg(self, "arch_type", val, str))
        if val not in ("coupling", "AR", "affine"):
            raise ValueError('NormalizingFlow arch_type must be "coupling", "AR", or "affine".')
        self._arch_type = val

    @property
    def num_stages(self):
        return self._num_stages

    @num_stages.setter
    def num_stages(self, val):
        if type(val) is not int:
            raise TypeError(format_type_err_msg(self, "num_stages", val, int))
        if val < 1:
            raise ValueError("NormalizingFlow num_stages %d must be greater than 0." % val)
        self._num_stages = val

    @property
    def num_layers(self):
        return self._num_layers

    @num_layers.setter
    def num_layers(self, val):
        if type(val) is not int:
            raise TypeError(format_type_err_msg(self, "num_layers", val, int))
        if val < 1:
            raise ValueError("NormalizingFlow num_layers arg %d must be greater than 0." % val)
        self._num_layers = val

    @property
    def num_units(self):
        return self._num_units

    @num_units.setter
    def num_units(self, val):
        if type(val) is not int:
            raise TypeError(format_type_err_msg(self, "num_units", val, int))
        if val < 1:
            raise ValueError("NormalizingFlow num_units %d must be greater than 0." % val)
        if val < 15:
            print("Warning: NormFlow.num_layers set to minimum of 15 (received %d)." % val)
            val = 15
        self._num_units = val

    def count_num_params(self):
        self.D_params = 0
        for b in self.bijectors:
            self.D_params += b.count_num_params()

    def _param_init(self):
        """xavier_normal_ on a (1, D_params) leaf tensor kept as a plain
        attribute (density_estimator.py:352-356)."""
        self.params = torch.nn.init.xavier_normal_(torch.zeros(1, self.D_params, requires_grad=True))
        return None

    def to(self, device):
        """Keep the unconditional weights resident on ``device`` (HBM)."""
        if not self.conditioner:
            self.params = self.params.detach().to(device).requires_grad_(True)
        return self

    # ---- sampling ------------------------------------------------------
    def __call__(self, N=100, params=None, freeze_bn=False):
        if not self.conditioner:
            return self.forward(self.params, N, freeze_bn=freeze_bn)
        return self.forward(params, N, freeze_bn=freeze_bn)

    def forward(self, params, N=100, freeze_bn=False, omega=None):
        """Sample ``N`` points per parameter row and their log-density
        (density_estimator.py:364-388). Returns ``(z (M,N,D) float32,
        log_q_z (M,N) float64)`` on the device of ``params``."""
        home = params.device
        pd = ops.to_device(params, torch.float32)
        self._check_params(pd)
        M = pd.size(0)
        if not (torch.is_grad_enabled() and pd.requires_grad):
            res = self._chain_sample(pd.detach(), M, N, omega, freeze_bn, home, src=params)
            if res is not None:
                return _to(res[0], home), _to(res[1], home)
        z, log_q = self._base(M, N, pd.device, omega)
        if torch.is_grad_enabled() and pd.requires_grad:
            z, log_q = self._forward_autograd(z, log_q, pd, freeze_bn)
        else:
            z, log_q = self._forward_plan(z, log_q, pd.detach(), freeze_bn, home, src=params)
        return _to(z, home), _to(log_q, home)

    def _base(self, M, N, device, omega):
        if omega is None:
            # one draw from numpy's global stream seeds the device Philox stream, so
            # np.random.seed(s) makes sampling reproducible as it does for the reference
            seed = int(np.random.randint(0, 2 ** 31 - 1)) * 2654435761 + 12345
            return ops.base_sample(M, N, self.D, seed, 0, device)
        if isinstance(omega, np.ndarray):
            omega = torch.from_numpy(np.ascontiguousarray(omega))
        if tuple(omega.shape) != (M, N, self.D):
            raise ValueError("omega must have shape %s, got %s" % ((M, N, self.D), tuple(omega.shape)))
        z = ops.to_device(omega, torch.float32).contiguous()
        return z, ops.base_logq(z)

    def _slices(self):
        """[(bijector, first parameter column, count)] in chain order."""
        out, idx = [], 0
        for b in self.bijectors:
            n = b.count_num_params()
            out.append((b, idx, n))
            idx += n
        return out

    def _use_tc(self, b, pd, z):
        mode = config.tc_precision()
        return (mode is not None and pd.shape[0] == 1 and z.dtype == torch.float32
                and b.name == "RealNVP" and z.shape[0] * z.shape[1] >= config.tc_min_rows()
                and ops.tc_supported(b.D, b.num_units, b.num_layers, mode))

    def _tc_train(self, b, pd, z):
        """The differentiable path of a coupling layer runs on tensor cores (forward ``tnf_coupling_tc``, backward
        ``tnf_coupling_tc_bwd``) when the weights are shared and the shape is compiled: in the bf16-conditioner mode, or
        in any tensor-core mode after ``config.set_training_backward("bf16")`` (fp32-parity forward, bf16 gradients);
        otherwise training stays on the exact CUDA-core kernels."""
        return (config.tc_backward_enabled() and pd.shape[0] == 1 and z.dtype == torch.float32
                and z.shape[0] * z.shape[1] >= config.tc_min_rows()
                and ops.tc_bwd_supported(b.D, b.num_units, b.num_layers)
                and ops.tc_supported(b.D, b.num_units, b.num_layers, config.tc_precision()))

    def _packed(self, b, idx, n, pd, src=None):
        """Packed tensor-core operand images of bijector ``b``'s weights (``pd[0, idx:idx+n]``).

        Cached per bijector, keyed on the CALLER's parameter tensor ``src``: the entry holds a strong reference to
        it (its address cannot be recycled by another tensor while the entry lives) and is valid only for that very
        object at the same in-place version.  A device copy made for this call (``pd`` when ``src`` lives on the
        host) is never a key: the caching allocator hands its address out again.  Without ``src`` the weights are
        repacked (one ~60 us kernel per layer)."""
        mode = config.tc_precision()
        if src is not None:
            hit = self._tc_cache.get(id(b))
            if (hit is not None and hit[0] is src and hit[1] == src._version and hit[2] == (idx, mode)
                    and hit[3].device == pd.device):
                return hit[3]
        packed = ops.tc_pack(pd[0, idx:idx + n], b.D, b.num_units, b.num_layers, b.transform_upper, precision=mode)
        if src is not None:
            self._tc_cache[id(b)] = (src, src._version, (idx, mode), packed)
        return packed

    def _check_params(self, pd):
        """The reference's ``params[:, idx:idx+n].view(...)`` raises on a short parameter matrix
        (bijectors.py:224-235); the kernels would read out of bounds instead."""
        if pd.dim() != 2 or pd.shape[1] < self.D_params:
            raise ValueError("NormFlow needs a (M, >=%d) parameter matrix, got %s" % (self.D_params, tuple(pd.shape)))

    def _can_fold(self, pd, z):
        """BatchNorm / Affine are folded into the next tensor-core coupling kernel when every coupling
        layer of the chain runs on that path (shared weights, fp32, bf16-conditioner mode)."""
        cps = [b for b in self.bijectors if b.name == "RealNVP"]
        return bool(cps) and all(self._use_tc(b, pd, z) for b in cps)

    # ---- whole-chain C-ABI calls (tnf_chain_logprob / tnf_chain_sample) -----------------------------------
    def _chain_pod(self, pd, z_like, src, sample, freeze_bn=False):
        """The chain as a ``tnf_bijector_t`` array (include/tnf.h) plus the objects that must stay alive during the
        call, or None when a bijector is outside the executor (MAF, user-defined bijectors)."""
        from . import _lib
        dev = pd.device
        arr = (_lib.Bijector * len(self.bijectors))()
        keep, bn_new = [], []
        for i, (b, idx, n) in enumerate(self._slices()):
            e = arr[i]
            e.param_offset = idx
            if b.name == "RealNVP":
                e.kind, e.num_layers, e.num_units, e.transform_upper = _lib.TNF_BIJ_REALNVP, b.num_layers, b.num_units, int(b.transform_upper)
                if self._use_tc(b, pd, z_like):
                    packed = self._packed(b, idx, n, pd, src)
                    keep.append(packed)
                    e.packed = packed.data_ptr()
                    if ops.kernel_timer is not None:      # bench.py: device time of every tensor-core coupling launch
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(); e1.record()          # creates the cudaEvent_t handles; re-recorded by the library
                        e.ev_start, e.ev_stop = e0.cuda_event, e1.cuda_event
                        ops.kernel_timer.append((e0, e1))
            elif b.name == "BatchNorm":
                e.kind, e.bn_eps = _lib.TNF_BIJ_BATCHNORM, float(b.eps)
                if sample and not freeze_bn:     # the call WRITES the batch statistics
                    mean = torch.empty(b.D, dtype=torch.float32, device=dev)
                    alpha = torch.empty(b.D, dtype=torch.float32, device=dev)
                    ld = torch.empty((), dtype=torch.float32, device=dev)
                    bn_new.append((b, mean, alpha, ld))
                else:
                    mean, alpha = b._state_on(dev, torch.float32)
                    ld = b._last_ld.to(device=dev, dtype=torch.float32)
                keep += [mean, alpha, ld]
                e.bn_mean, e.bn_alpha, e.bn_log_det = mean.data_ptr(), alpha.data_ptr(), ld.data_ptr()
            elif b.name == "Affine":
                e.kind = _lib.TNF_BIJ_AFFINE
            elif b.name == "ToInterval":
                e.kind = _lib.TNF_BIJ_TOINTERVAL
                c = b._consts(dev)
                keep.append(c)
                e.consts = c.data_ptr()
            elif b.name == "ToSimplex" and sample:
                e.kind, e.num_units = _lib.TNF_BIJ_TOSIMPLEX, b.D
            else:
                return None
        return arr, keep, bn_new

    def _chain_logprob(self, zd, pd, src=None, use_tc=None):
        """log_prob of device samples through ONE C-ABI call; None when the chain needs the per-bijector plan."""
        from . import _lib
        if zd.dtype != torch.float32 or pd.dtype != torch.float32 or not config.chain_abi():
            return None
        M, N, D = zd.shape
        z_like = zd if use_tc is None else _Rows(M, (1 << 40) if use_tc else 0, zd.dtype)
        pod = self._chain_pod(pd, z_like, src, sample=False)
        if pod is None:
            return None
        arr, keep, _ = pod
        Mp = pd.shape[0]
        if Mp != M and Mp != 1:
            raise ValueError("params has %d rows but z has M=%d" % (Mp, M))
        lp = torch.empty((M, N), dtype=torch.float32, device=zd.device)
        nbytes = _lib.lib().tnf_chain_workspace_bytes(M, N, D)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=zd.device)
        prec = ops.TC_PRECISION.get(config.tc_precision(), 0)
        rc = _lib.lib().tnf_chain_logprob(arr, len(self.bijectors), zd.data_ptr(), pd.data_ptr(),
                                          pd.stride(0) if Mp > 1 else 0, M, N, D, prec, lp.data_ptr(), ws.data_ptr(),
                                          nbytes, ops._stream())
        _lib.check(rc, "tnf_chain_logprob")
        return lp

    def _chain_sample(self, pd, M, N, omega, freeze_bn, home, src=None):
        """(z, log_q) through ONE C-ABI call; None when the chain needs the per-bijector plan."""
        from . import _lib, dist
        if pd.dtype != torch.float32 or not config.chain_abi():
            return None
        dev = pd.device
        D = self.D
        pod = self._chain_pod(pd, _Rows(M, N, torch.float32), src, sample=True, freeze_bn=freeze_bn)
        if pod is None:
            return None
        arr, keep, bn_new = pod
        Mp = pd.shape[0]
        d_out = D + 1 if self.bijectors and self.bijectors[-1].name == "ToSimplex" else D
        z = torch.empty((M, N, d_out), dtype=torch.float32, device=dev)
        log_q = torch.empty((M, N), dtype=torch.float64, device=dev)
        seed = 0
        om_ptr = 0
        if omega is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1)) * 2654435761 + 12345
        else:
            if isinstance(omega, np.ndarray):
                omega = torch.from_numpy(np.ascontiguousarray(omega))
            if tuple(omega.shape) != (M, N, D):
                raise ValueError("omega must have shape %s, got %s" % ((M, N, D), tuple(omega.shape)))
            omega = ops.to_device(omega, torch.float32).contiguous()
            om_ptr = omega.data_ptr()
        nbytes = _lib.lib().tnf_chain_workspace_bytes(M, N, D)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        stats = None
        cb = _lib.ALLREDUCE_FN(0)
        if dist.is_enabled() and dist.world_size() > 1 and not freeze_bn:
            stats = torch.empty(2 * D + 1, dtype=torch.float64, device=dev)

            def _hook(_ptr, _count, _user, _stats=stats):
                try:
                    dist.allreduce_stats(_stats)       # in place, enqueued on the current stream
                    return 0
                except Exception:                      # never unwind through the C frame
                    return 1
            cb = _lib.ALLREDUCE_FN(_hook)
        prec = ops.TC_PRECISION.get(config.tc_precision(), 0)
        peer = None
        if stats is not None:      # statistics over NVLink peer memory inside the fold kernels, when set up (dist.py)
            peer = dist.peer_struct(sum(1 for b in self.bijectors if b.name == "BatchNorm"))
        import ctypes
        rc = _lib.lib().tnf_chain_sample(arr, len(self.bijectors), pd.data_ptr(), pd.stride(0) if Mp > 1 else 0, M, N, D,
                                         prec, om_ptr, seed & (2 ** 64 - 1), 0, int(bool(freeze_bn)), cb, None,
                                         stats.data_ptr() if stats is not None else None,
                                         ctypes.byref(peer) if peer is not None else None, z.data_ptr(), log_q.data_ptr(),
                                         ws.data_ptr(), nbytes, ops._stream())
        _lib.check(rc, "tnf_chain_sample")
        for (b, mean, alpha, ld) in bn_new:
            b._set_state(mean, alpha, ld, home)
        return z, log_q

    def _forward_plan(self, z, log_q, pd, freeze_bn, home=None, src=None):
        M, N, D = z.shape
        Mp = pd.shape[0]
        ld_acc = torch.zeros((M, N), dtype=torch.float32, device=z.device)
        scal = torch.zeros(Mp, dtype=torch.float32, device=z.device)
        fold = self._can_fold(pd, z)
        pend = None     # per-column map (scale, shift) owed to z; consumed by the next coupling kernel
        stats = None    # column statistics of z produced by the previous coupling kernel itself
        slices = self._slices()
        for i, (b, idx, n) in enumerate(slices):
            if b.name == "RealNVP":
                if self._use_tc(b, pd, z):
                    nxt = slices[i + 1][0].name if i + 1 < len(slices) else None
                    mode = config.tc_precision()
                    fuse_stats = nxt == "BatchNorm" and not freeze_bn      # every tensor-core kernel emits them
                    res = ops.coupling_tc(z, self._packed(b, idx, n, pd, src), b.D, b.num_units, b.num_layers,
                                          b.transform_upper, TNF_FORWARD, ld=ld_acc, accum=TNF_LD_ADD,
                                          pre_scale=pend[0] if pend else None, pre_shift=pend[1] if pend else None,
                                          want_stats=fuse_stats, precision=mode)
                    z, stats = res[0], (res[2] if fuse_stats else None)
                    pend = None
                else:
                    z, _ = ops.coupling(z, pd[:, idx:idx + n], b.D, b.num_units, b.num_layers, b.transform_upper,
                                        TNF_FORWARD, ld=ld_acc, accum=TNF_LD_ADD)
                    stats = None
            elif b.name == "BatchNorm":
                if pend is not None:       # statistics are taken on the materialised tensor
                    z, pend, stats = ops.colaffine(z, pend, D), None, None
                if freeze_bn:
                    mean, alpha = b._state_on(z.device, z.dtype)
                    ld = b._last_ld.to(device=z.device, dtype=z.dtype)
                else:
                    from .bijectors import _reduce_stats
                    sums = _reduce_stats(stats if stats is not None else ops.colstats(z, D))
                    mean, alpha, ld = ops.bn_finalize(sums, D, b.eps, z.dtype)
                    b._set_state(mean, alpha, ld, home)
                stats = None
                if fold:
                    pend = ops.fold_colaffine(pend, ops.FOLD_BN_FWD, mean, alpha, D)
                else:
                    z = ops.bn_apply(z, mean, alpha, D, TNF_FORWARD)
                ops.accum_bcast(scal, ld.reshape(1), Mp)
            elif b.name == "Affine":
                if fold:
                    pend = ops.fold_colaffine(pend, ops.FOLD_AFF_FWD, pd[0, idx:idx + D].contiguous(),
                                              pd[0, idx + D:idx + 2 * D].contiguous(), D, ld_accum=scal)
                else:
                    z, ld = ops.affine(z, pd[:, idx:idx + n], D, TNF_FORWARD)
                    ops.accum_bcast(scal, ld, 1)
            else:
                if pend is not None:
                    z, pend = ops.colaffine(z, pend, D), None
                if b.name == "ToInterval":
                    z, _ = ops.tointerval(z, b._consts(z.device), D, TNF_FORWARD, ld=ld_acc, accum=TNF_LD_ADD)
                elif b.name == "ToSimplex":
                    z, _ = ops.tosimplex(z, b.D, ld=ld_acc, accum=TNF_LD_ADD)
                else:  # MAF or a user-defined bijector: public protocol, log-det folded in by kernel
                    z, ld = b(z, pd[:, idx:idx + n]) if n > 0 else b(z)
                    ld = ld.detach().to(torch.float32)
                    if ld.numel() == M * N:
                        ops.accum_bcast(ld_acc, ld.contiguous(), 1)
                    else:
                        ops.accum_bcast(scal, ld.reshape(-1).contiguous(), Mp if ld.numel() == 1 else 1)
                    z = z.detach()
        if pend is not None:
            z = ops.colaffine(z, pend, D)
        ops.finish_logq(log_q, ld_acc, scal, N if Mp == M and M > 1 else M * N)
        return z, log_q

    def _forward_autograd(self, z, log_q, pd, freeze_bn):
        """Differentiable sample path: per-bijector autograd functions
        (density_estimator.py:374-387 verbatim control flow)."""
        for (b, idx, n) in self._slices():
            if b.name == "BatchNorm":
                z, log_det = b(z, use_last=freeze_bn)
            elif n > 0:
                z, log_det = b(z, pd[:, idx:idx + n])
            else:
                z, log_det = b(z)
            log_q = log_q - log_det
        return z, log_q

    # ---- density -------------------------------------------------------
    def inverse_and_log_det(self, z, params):
        """Run the chain backwards (density_estimator.py:390-406). Returns
        ``(z0, sum_log_det (M,N))``."""
        home = z.device
        zd = ops.to_device(z if z.dtype in (torch.float32, torch.float64) else z.float())
        pd = ops.to_device(params, zd.dtype)
        self._check_params(pd)
        if torch.is_grad_enabled() and (pd.requires_grad or zd.requires_grad):
            z0, sld = self._inverse_autograd(zd, pd)
        else:
            z0, ld_acc, scal, div = self._inverse_plan(zd.detach().contiguous(), pd.detach(), src=params)
            sld = ops.accum_bcast(ld_acc, scal, div)
        return _to(z0, home), _to(sld, home)

    def _inverse_plan(self, z, pd, src=None, use_tc=None):
        """``use_tc``: tensor-core eligibility decided by the caller on the WHOLE batch (the host pipeline hands in
        row chunks; a chunk must not pick a different precision path than the one-shot call would)."""
        M, N, D = z.shape
        Mp = pd.shape[0]
        ld_acc = torch.zeros((M, N), dtype=z.dtype, device=z.device)
        scal = torch.zeros(Mp, dtype=z.dtype, device=z.device)
        fold = self._can_fold(pd, z) if use_tc is None else use_tc
        pend = None
        for (b, idx, n) in reversed(self._slices()):
            if b.name == "RealNVP":
                if (self._use_tc(b, pd, z) if use_tc is None else use_tc):
                    z, _ = ops.coupling_tc(z, self._packed(b, idx, n, pd, src), b.D, b.num_units, b.num_layers,
                                           b.transform_upper, TNF_INVERSE, ld=ld_acc, accum=TNF_LD_ADD,
                                           pre_scale=pend[0] if pend else None, pre_shift=pend[1] if pend else None,
                                           precision=config.tc_precision())
                    pend = None
                else:
                    z, _ = ops.coupling(z, pd[:, idx:idx + n], b.D, b.num_units, b.num_layers, b.transform_upper,
                                        TNF_INVERSE, ld=ld_acc, accum=TNF_LD_ADD)
            elif b.name == "BatchNorm":
                mean, alpha = b._state_on(z.device, z.dtype)
                if fold:
                    pend = ops.fold_colaffine(pend, ops.FOLD_BN_INV, mean, alpha, D)
                else:
                    z = ops.bn_apply(z, mean, alpha, D, TNF_INVERSE)
                ops.accum_bcast(scal, b._last_ld.to(device=z.device, dtype=z.dtype).reshape(1), Mp)
            elif b.name == "Affine":
                if fold:
                    pend = ops.fold_colaffine(pend, ops.FOLD_AFF_INV, pd[0, idx:idx + D].contiguous(),
                                              pd[0, idx + D:idx + 2 * D].contiguous(), D, ld_accum=scal)
                else:
                    z, ld = ops.affine(z, pd[:, idx:idx + n], D, TNF_INVERSE)
                    ops.accum_bcast(scal, ld, 1)
            else:
                if pend is not None:
                    z, pend = ops.colaffine(z, pend, D), None
                if b.name == "ToInterval":
                    z, _ = ops.tointerval(z, b._consts(z.device), D, TNF_INVERSE, ld=ld_acc, accum=TNF_LD_ADD)
                else:
                    if n > 0:
                        z, ld = b.inverse_and_log_det(z, pd[:, idx:idx + n])
                    else:
                        z, ld = b.inverse_and_log_det(z)   # ToSimplex: TypeError, as in the reference
                    ld = ld.detach().to(z.dtype)
                    if ld.numel() == M * N:
                        ops.accum_bcast(ld_acc, ld.contiguous(), 1)
                    else:
                        ops.accum_bcast(scal, ld.reshape(-1).contiguous(), Mp if ld.numel() == 1 else 1)
                    z = z.detach()
        if pend is not None:
            z = ops.colaffine(z, pend, D)
        return z, ld_acc, scal, (N if Mp == M and M > 1 else M * N)

    def _chain_grad_ok(self):
        """The differentiable ``log_prob`` runs as ONE autograd node (``_ChainLogProbFn``) when every bijector is one the
        node knows how to differentiate; otherwise bijector by bijector."""
        return all(b.name in ("RealNVP", "BatchNorm", "Affine", "ToInterval") for b in self.bijectors)

    def _inverse_autograd(self, z, pd):
        sum_log_det = torch.zeros((z.shape[0], z.shape[1]), dtype=z.dtype, device=z.device)
        for (b, idx, n) in reversed(self._slices()):
            if n > 0:
                z, log_det = b.inverse_and_log_det(z, pd[:, idx:idx + n])
            else:
                z, log_det = b.inverse_and_log_det(z)
            sum_log_det = sum_log_det + log_det
        return z, sum_log_det

    def _log_prob_host_pipelined(self, z, pd, src=None):
        """``log_prob`` of host samples ``z (1, N, D)`` in row chunks: the host->device copy of chunk i+1 (copy stream)
        overlaps the inverse chain of chunk i (current stream).  Exact: no bijector couples samples in this direction
        (BatchNorm.inverse_and_log_det uses its stored statistics, bijectors.py:420-426)."""
        M, N, D = z.shape
        n_chunks = config.host_pipeline_chunks()
        step = -(-N // n_chunks)
        step = (step + 127) // 128 * 128                      # whole 128-row tiles per chunk
        dev = torch.device("cuda", torch.cuda.current_device())
        compute, copy = torch.cuda.current_stream(), _copy_stream()
        zd = torch.empty((M, N, D), dtype=z.dtype, device=dev)
        lp = torch.empty((M, N), dtype=z.dtype, device=dev)
        ready = torch.cuda.Event()
        ready.record(compute)                                  # zd's block may still be in use by earlier kernels
        copy.wait_event(ready)
        bounds = [(lo, min(N, lo + step)) for lo in range(0, N, step)]
        use_tc = self._can_fold(pd, z)        # decided once, on the full batch

        def issue_copy(lo, hi):
            with torch.cuda.stream(copy):
                zd[:, lo:hi].copy_(z[:, lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            return ev

        ev = issue_copy(*bounds[0])
        for i, (lo, hi) in enumerate(bounds):
            compute.wait_event(ev)
            part = self._chain_logprob(zd[:, lo:hi], pd, src=src, use_tc=use_tc)
            if part is None:
                z0, ld_acc, scal, div = self._inverse_plan(zd[:, lo:hi], pd, src=src, use_tc=use_tc)
                part = ops.base_logprob(z0, ld_acc, scal, div)
            lp[:, lo:hi] = part
            if i + 1 < len(bounds):                            # queued after chunk i's kernels: a pageable source blocks
                ev = issue_copy(*bounds[i + 1])                # the host here while the GPU works on chunk i
        return lp

    def log_prob(self, z, params=None):
        """log q(z) = log N(z0; 0, I) - sum_log_det (density_estimator.py:408-416)."""
        if not self.conditioner:
            params = self.params
        home = z.device
        if z.dtype not in (torch.float32, torch.float64):
            z = z.float()
        needs_grad = torch.is_grad_enabled() and (params.requires_grad or z.requires_grad)
        if (not z.is_cuda and not needs_grad and z.dim() == 3 and z.shape[0] == 1 and params.shape[0] == 1
                and z.shape[1] >= config.host_pipeline_min_rows() and config.host_pipeline_chunks() > 1):
            ops.require_cuda()
            pd = ops.to_device(params.detach(), z.dtype)
            self._check_params(pd)
            return _to(self._log_prob_host_pipelined(z.detach().contiguous(), pd, src=params), home)
        zd = ops.to_device(z)
        pd = ops.to_device(params, zd.dtype)
        self._check_params(pd)
        if torch.is_grad_enabled() and (pd.requires_grad or zd.requires_grad):
            if self._chain_grad_ok():
                lp = _ChainLogProbFn.apply(zd.contiguous(), pd, self)
            else:
                z0, sld = self._inverse_autograd(zd, pd)
                lp = _BaseLogProbFn.apply(z0) - sld
        else:
            lp = self._chain_logprob(zd.detach().contiguous(), pd.detach(), src=params)
            if lp is None:
                z0, ld_acc, scal, div = self._inverse_plan(zd.detach().contiguous(), pd.detach(), src=params)
                lp = ops.base_logprob(z0, ld_acc, scal, div)
        return _to(lp, home)


class _Rows(object):
    """Stand-in for a z tensor in ``_use_tc`` (shape and dtype are all it reads)."""

    def __init__(self, M, N, dtype):
        self.shape, self.dtype = (M, N, 0), dtype


_copy_streams = {}


def _copy_stream():
    dev = torch.cuda.current_device()
    if dev not in _copy_streams:
        _copy_streams[dev] = torch.cuda.Stream(device=dev)
    return _copy_streams[dev]


class _ChainLogProbFn(torch.autograd.Function):
    """``log_prob`` of a whole chain as one autograd node (density_estimator.py:390-416 differentiated).

    The per-bijector autograd path slices ``params`` per bijector, so autograd zero-fills, copies and adds one
    ``(M, D_params)`` gradient tensor per bijector (at C4 that glue was 20 % of the training step).  Here the forward
    keeps every bijector's input, and the backward walks the chain once, each ``*_bwd`` kernel accumulating straight
    into ITS columns of one gradient buffer (pointer + row stride, as the kernels take their parameters)."""

    @staticmethod
    def forward(ctx, z, pd, nf):
        M, N, D = z.shape
        ins = []
        ld_acc = torch.zeros((M, N), dtype=z.dtype, device=z.device)
        scal = torch.zeros(pd.shape[0], dtype=z.dtype, device=z.device)
        pdd = pd.detach()
        cur = z.detach()
        order = list(reversed(nf._slices()))             # execution order of the inverse chain
        pre = None                                       # (alpha, mean) of a BatchNorm folded into the NEXT coupling layer
        folded = {}                                      # position in `order` of a coupling layer -> its folded pre-affine
        tc_pos = set()                                   # positions whose coupling layer ran (and will be differentiated) on tensor cores
        for k, (b, idx, n) in enumerate(order):
            ins.append(cur)
            if b.name == "RealNVP" and nf._tc_train(b, pdd, cur):
                # bf16-conditioner mode, shared weights: forward AND backward of the layer on tensor cores; a BatchNorm
                # right before it is applied on load (z alpha + mean with its remembered statistics: constants)
                ps, pb = pre if pre is not None else (None, None)
                tc_pos.add(k)
                if pre is not None:
                    folded[k] = pre
                    pre = None
                cur, _ = ops.coupling_tc(cur, nf._packed(b, idx, n, pdd, src=pd), b.D, b.num_units, b.num_layers,
                                         b.transform_upper, TNF_INVERSE, ld=ld_acc, accum=TNF_LD_ADD, pre_scale=ps,
                                         pre_shift=pb, precision=config.tc_precision())
            elif b.name == "RealNVP":
                cur, _ = ops.coupling(cur, pdd[:, idx:idx + n], b.D, b.num_units, b.num_layers, b.transform_upper, TNF_INVERSE,
                                      ld=ld_acc, accum=TNF_LD_ADD)
            elif b.name == "BatchNorm":
                mean, alpha = b._state_on(z.device, z.dtype)
                nxt = order[k + 1][0] if k + 1 < len(order) else None
                if (nxt is not None and nxt.name == "RealNVP" and nf._tc_train(nxt, pdd, cur)
                        and mean.numel() == D and alpha.numel() == D):       # (identity statistics before the first forward are scalars)
                    pre = (alpha.reshape(-1).contiguous(), mean.reshape(-1).contiguous())
                else:
                    cur = ops.bn_apply(cur, mean, alpha, D, TNF_INVERSE)
                ops.accum_bcast(scal, b._last_ld.to(device=z.device, dtype=z.dtype).reshape(1), pd.shape[0])
            elif b.name == "Affine":
                cur, ld = ops.affine(cur, pdd[:, idx:idx + n], D, TNF_INVERSE)
                ops.accum_bcast(scal, ld, 1)
            else:   # ToInterval
                cur, _ = ops.tointerval(cur, b._consts(z.device), D, TNF_INVERSE, ld=ld_acc, accum=TNF_LD_ADD)
        lp = ops.base_logprob(cur, ld_acc, scal, N if pd.shape[0] == M and M > 1 else M * N)
        ctx.nf, ctx.ins, ctx.z0, ctx.folded, ctx.tc_pos = nf, ins, cur, folded, tc_pos
        ctx.save_for_backward(pd)
        ctx.need = (z.requires_grad, pd.requires_grad)
        return lp

    @staticmethod
    def backward(ctx, g_lp):
        (pd,) = ctx.saved_tensors
        nf, ins, z0 = ctx.nf, ctx.ins, ctx.z0
        M, N, D = z0.shape
        Mp = pd.shape[0]
        g_lp = g_lp.contiguous()
        pdd = pd.detach()
        # one row per context: the coupling backward WRITES its columns (no zero-fill, no read-modify-write of the
        # (M, D_params) matrix); otherwise a zero-filled buffer that the kernels accumulate into
        ow = Mp == M and M > 1 and N <= 32
        g_params = (torch.empty if ow else torch.zeros)(pd.shape, dtype=pd.dtype, device=pd.device)
        if ow and pd.shape[1] > nf.D_params:
            g_params[:, nf.D_params:].zero_()            # trailing extra columns (legal, ignored) get a zero gradient
        g_z = ops.base_logprob_bwd(z0, g_lp)               # d log N(z0) / d z0 = -z0
        g_ld = -g_lp                                       # log_prob = log N(z0) - sum of log-dets
        g_ld_rows = g_ld.sum(dim=1) if Mp == M and M > 1 else g_ld.sum().reshape(1)     # Affine: log-det per parameter row
        slices = nf._slices()                              # chain order = the reverse of the forward's execution order
        folded = ctx.folded
        for k, (b, idx, n) in enumerate(slices):
            pos = len(slices) - 1 - k                    # this bijector's position in the forward's execution order
            z_in = ins[pos]
            if b.name == "RealNVP" and pos in ctx.tc_pos:      # as decided in the forward
                packed_b = ops.tc_bwd_pack(pdd[0, idx:idx + n], b.D, b.num_units, b.num_layers, b.transform_upper)
                ps, pb = folded.get(pos, (None, None))
                g_z = ops.coupling_tc_bwd(z_in, packed_b, g_z, g_ld, g_params[0, idx:idx + n], b.D, b.num_units,
                                          b.num_layers, b.transform_upper, TNF_INVERSE, pre_scale=ps, pre_shift=pb)
            elif b.name == "RealNVP":
                g_z = ops.coupling_bwd(z_in, pdd[:, idx:idx + n], g_z, g_ld, g_params[:, idx:idx + n], b.D, b.num_units,
                                       b.num_layers, b.transform_upper, TNF_INVERSE, overwrite=ow)
            elif b.name == "BatchNorm":                    # remembered statistics are constants: z alpha + mean
                if (pos + 1) in folded:                    # folded into the coupling layer after it: that kernel scaled g_z
                    continue
                _, alpha = b._state_on(z0.device, z0.dtype)
                g_z = ops.bn_apply(g_z, torch.zeros_like(alpha), alpha, D, TNF_INVERSE)
            elif b.name == "Affine":
                if ow:
                    g_params[:, idx:idx + n].zero_()       # 2 D columns: the Affine backward accumulates
                g_z = ops.affine_bwd(z_in, pdd[:, idx:idx + n], g_z, g_ld_rows, g_params[:, idx:idx + n], D, TNF_INVERSE)
            else:
                g_z = ops.tointerval_bwd(z_in, b._consts(z0.device), g_z, g_ld, D, TNF_INVERSE)
        ctx.ins = ctx.z0 = None
        return (g_z if ctx.need[0] else None), (g_params if ctx.need[1] else None), None


class _BaseLogProbFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z):
        ctx.save_for_backward(z)
        return ops.base_logprob(z)

    @staticmethod
    def backward(ctx, g):
        (z,) = ctx.saved_tensors
        return ops.base_logprob_bwd(z, g)


def _to(t, device):
    return ops.to_like(t, device)
