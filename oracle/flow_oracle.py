"""CPU oracle for the torch_nf bijector-chain hot path.  TEST INFRASTRUCTURE ONLY.

This module is a functional CPU restatement of the reference algorithm
(srbittner/torch_nf).  It exists so that the CUDA path can be checked; it is
never the thing shipped or measured.  Only ``tests/``, ``__graft_entry__.smoke``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
Nothing under ``torch_nf_b200/`` imports it.

Arithmetic library: the reference's arithmetic *is* torch (``torch.matmul``,
``torch.tanh``, ``torch.exp`` ... on CPU tensors; ``setup.py:10`` lists bare
``torch``, i.e. unpinned; this image has torch 2.11.0).  The restatement
therefore uses the same CPU torch primitives in the dtype it is handed
(float32 like the reference's ``.float()`` tensors, or float64 for a
high-precision truth), so that it rounds like the reference and costs what the
reference costs when it is timed as the CPU baseline.

Parity pin: ``tests/golden/*.npz`` were produced by importing the UNMODIFIED
reference from ``/root/reference`` (``tests/golden/make_golden.py``) and
``tests/test_oracle_golden.py`` checks every function here against them.

Stateless by design: where the reference keeps state on the object
(``BatchNorm.__last_mean/__last_alpha``) the oracle takes and returns it.

Citations ``file:line`` are relative to the reference repository root.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------
# RealNVP coupling layer               (torch_nf/bijectors.py:145-262)
# ---------------------------------------------------------------------------
def coupling_dims(D, transform_upper):
    """(D_in, D_out) of the conditioner.  torch_nf/bijectors.py:157-165."""
    h = D // 2
    d_in, d_out = h, h
    if D % 2 == 1:
        if transform_upper:
            d_out += 1
        else:
            d_in += 1
    return d_in, d_out


def coupling_num_params(D, num_layers, num_units, transform_upper):
    """torch_nf/bijectors.py:244-262."""
    d_in, d_out = coupling_dims(D, transform_upper)
    U = num_units
    return 2 * (d_in * U + d_out * U + d_out + U + (num_layers - 1) * (U + 1) * U)


def _conditioner(z1, params, d_in, d_out, num_layers, U):
    """The paired shift/scale MLPs.  Packing per layer, read from the front of
    ``params``: W_t[K*J] (row = input), W_s[K*J], b_t[J], b_s[J]
    (torch_nf/bijectors.py:224-242); layer sizes d_in->U (tanh),
    U->U (tanh) x (L-1), U->d_out (linear) (torch_nf/bijectors.py:168-171)."""
    M = params.shape[0]
    sizes = [(d_in, U, True)] + [(U, U, True)] * (num_layers - 1) + [(U, d_out, False)]
    t = s = z1
    off = 0
    for (K, J, act) in sizes:
        Wt = params[:, off:off + K * J].reshape(M, K, J); off += K * J
        Ws = params[:, off:off + K * J].reshape(M, K, J); off += K * J
        bt = params[:, off:off + J].reshape(M, 1, J); off += J
        bs = params[:, off:off + J].reshape(M, 1, J); off += J
        t = torch.matmul(t, Wt) + bt
        s = torch.matmul(s, Ws) + bs
        if act:
            t = torch.tanh(t)
            s = torch.tanh(s)
    return t, s


def _split(z, D, transform_upper):
    h = D // 2
    if transform_upper:
        return z[:, :, :h], z[:, :, h:]          # z1 (conditions), z2 (transformed)
    return z[:, :, h:], z[:, :, :h]


def _join(z1, z2, transform_upper):
    return torch.cat([z1, z2], dim=2) if transform_upper else torch.cat([z2, z1], dim=2)


def coupling_forward(z, params, D, num_layers, num_units, transform_upper=True):
    """z2' = t + z2*exp(s); log_det = sum(s).  torch_nf/bijectors.py:145-181."""
    d_in, d_out = coupling_dims(D, transform_upper)
    z1, z2 = _split(z, D, transform_upper)
    t, s = _conditioner(z1, params, d_in, d_out, num_layers, num_units)
    z2 = t + z2 * torch.exp(s)
    return _join(z1, z2, transform_upper), torch.sum(s, dim=2)


def coupling_inverse(z, params, D, num_layers, num_units, transform_upper=True):
    """z2' = (z2 - t)/exp(s); returns +sum(s).  torch_nf/bijectors.py:183-206."""
    d_in, d_out = coupling_dims(D, transform_upper)
    z1, z2 = _split(z, D, transform_upper)
    t, s = _conditioner(z1, params, d_in, d_out, num_layers, num_units)
    z2 = (z2 - t) / torch.exp(s)
    return _join(z1, z2, transform_upper), torch.sum(s, dim=2)


# ---------------------------------------------------------------------------
# Affine                                 (torch_nf/bijectors.py:265-318)
# ---------------------------------------------------------------------------
def affine_forward(z, params, D):
    alpha, shift = params[:, :D], params[:, D:2 * D]
    z = torch.exp(alpha)[:, None, :] * z + shift[:, None, :]
    return z, torch.sum(alpha, dim=1, keepdim=True)


def affine_inverse(z, params, D):
    alpha, shift = params[:, :D], params[:, D:2 * D]
    z = (z - shift[:, None, :]) / torch.exp(alpha)[:, None, :]
    return z, torch.sum(alpha, dim=1, keepdim=True)


# ---------------------------------------------------------------------------
# BatchNorm                              (torch_nf/bijectors.py:321-426)
# ---------------------------------------------------------------------------
def batchnorm_forward(z, eps=1e-5, use_last=False, last_mean=None, last_alpha=None):
    """Returns (z_norm, log_det, mean, alpha).

    The reference obtains ``alpha`` and ``mean`` through ``BatchNorm1d`` plus a
    ratio of variances (torch_nf/bijectors.py:402-410); algebraically that is
    ``alpha = sqrt(biased_var + eps)``, ``mean = batch mean`` over the
    flattened (M*N, D) batch, which is what is restated here.
    ``log_det = -sum(log alpha)`` is a scalar (torch_nf/bijectors.py:417)."""
    D = z.shape[-1]
    if use_last:
        mean, alpha = last_mean, last_alpha
    else:
        zv = z.reshape(-1, D)
        mean = zv.mean(dim=0)
        alpha = torch.sqrt(zv.var(dim=0, unbiased=False) + eps)
    z_norm = (z - mean) / alpha
    return z_norm, -torch.sum(torch.log(alpha)), mean, alpha


def batchnorm_inverse(z, last_mean, last_alpha):
    """torch_nf/bijectors.py:420-426 (two separate roundings: mul then add)."""
    z = z * last_alpha
    z = z + last_mean
    return z, -torch.sum(torch.log(last_alpha))


# ---------------------------------------------------------------------------
# ToInterval                             (torch_nf/bijectors.py:429-557)
# ---------------------------------------------------------------------------
_TI_EPS = 1e-12


def tointerval_consts(lb, ub):
    """Per-dimension mode constants.  torch_nf/bijectors.py:454-480.  The
    reference stores them as float32 tensors regardless of the input dtype."""
    lb = np.asarray(lb, dtype=np.float64)
    ub = np.asarray(ub, dtype=np.float64)
    D = lb.shape[0]
    tanh_flg = np.zeros(D); sp_flg = np.zeros(D)
    tanh_m = np.ones(D); tanh_c = np.zeros(D)
    sp_m = np.ones(D); sp_c = np.zeros(D)
    for i in range(D):
        has_lb = not np.isneginf(lb[i])
        has_ub = not np.isposinf(ub[i])
        if has_lb and has_ub:
            tanh_flg[i] = 1
            tanh_m[i] = (ub[i] - lb[i]) / 2.0
            tanh_c[i] = (ub[i] + lb[i]) / 2.0
        elif has_lb:
            sp_flg[i] = 1; sp_m[i] = 1.0; sp_c[i] = lb[i]
        elif has_ub:
            sp_flg[i] = 1; sp_m[i] = -1.0; sp_c[i] = ub[i]
    f32 = lambda a: torch.tensor(a).float()[None, None, :]
    return dict(tanh_flg=f32(tanh_flg), sp_flg=f32(sp_flg), tanh_m=f32(tanh_m),
                tanh_c=f32(tanh_c), sp_m=f32(sp_m), sp_c=f32(sp_c))


def _ti_tanh_ldj(c, z):
    th = torch.tanh(z)
    return torch.sum(c["tanh_flg"] * (torch.log(c["tanh_m"]) + torch.log(1.0 - th ** 2 + _TI_EPS)), dim=2)


def tointerval_forward(z, lb, ub):
    """torch_nf/bijectors.py:509-527 (mask blends, tanh stage then softplus stage)."""
    c = tointerval_consts(lb, ub)
    tanh_ldj = _ti_tanh_ldj(c, z)
    out = c["tanh_m"] * torch.tanh(z) + c["tanh_c"]
    z = c["tanh_flg"] * out + (1 - c["tanh_flg"]) * z
    out = c["sp_m"] * F.softplus(z) + c["sp_c"]
    sp_ldj = torch.sum(c["sp_flg"] * F.logsigmoid(z), dim=2)
    z = c["sp_flg"] * out + (1 - c["sp_flg"]) * z
    return z, tanh_ldj + sp_ldj


def tointerval_inverse(z, lb, ub):
    """torch_nf/bijectors.py:529-557; returns the forward ldj at the pre-image."""
    c = tointerval_consts(lb, ub)
    sp_inv = torch.log(torch.exp(c["sp_flg"] * (z - c["sp_c"]) / c["sp_m"]) - 1 + _TI_EPS)
    z = c["sp_flg"] * sp_inv + (1 - c["sp_flg"]) * z
    sp_ldj = torch.sum(c["sp_flg"] * F.logsigmoid(z), dim=2)
    x = c["tanh_flg"] * (z - c["tanh_c"]) / c["tanh_m"]
    tanh_inv = 0.5 * (torch.log(1 + x + _TI_EPS) - torch.log(1 - x + _TI_EPS))
    z = c["tanh_flg"] * tanh_inv + (1 - c["tanh_flg"]) * z
    return z, _ti_tanh_ldj(c, z) + sp_ldj


# ---------------------------------------------------------------------------
# ToSimplex                              (torch_nf/bijectors.py:560-594)
# ---------------------------------------------------------------------------
def tosimplex_forward(z, D):
    """(M,N,D)->(M,N,D+1); ``D`` is the bijector's D attribute used in the
    log-det (torch_nf/bijectors.py:580-589), which tests construct with a
    value one larger than the input width (tests/test_bijectors.py:350-358)."""
    ex = torch.exp(z)
    sum_ex = torch.sum(ex, dim=2)
    den = sum_ex + 1.0
    log_det = torch.log(1.0 - (sum_ex / den) + 1e-10) - D * torch.log(den) + torch.sum(z, dim=2)
    z = torch.cat((ex / den[:, :, None], 1.0 / den[:, :, None]), dim=2)
    return z, log_det


# ---------------------------------------------------------------------------
# NormFlow chain                         (torch_nf/density_estimator.py:240-421)
# ---------------------------------------------------------------------------
def build_chain(D, arch_type="coupling", num_stages=1, num_layers=2, num_units=15,
                support=None):
    """Bijector list as plain dicts, in the reference's order
    (torch_nf/density_estimator.py:260-282).  ``support`` is None,
    ("ToInterval", lb, ub) or ("ToSimplex", D_attr).  Clamps on num_layers /
    num_units follow torch_nf/bijectors.py:110-131, density_estimator.py:344-348."""
    num_units = max(int(num_units), 15)
    chain = []
    if arch_type == "coupling":
        L = min(int(num_layers), 5)
        U = min(num_units, 1000)
        for _ in range(num_stages):
            chain.append(dict(kind="RealNVP", L=L, U=U, upper=True))
            chain.append(dict(kind="BatchNorm", eps=1e-5))
            chain.append(dict(kind="RealNVP", L=L, U=U, upper=False))
            chain.append(dict(kind="BatchNorm", eps=1e-5))
            chain.append(dict(kind="Affine"))
    elif arch_type == "affine":
        chain.append(dict(kind="Affine"))
    else:
        raise ValueError("oracle covers arch_type 'coupling' and 'affine'")
    if support is not None:
        if support[0] == "ToInterval":
            chain.append(dict(kind="ToInterval", lb=support[1], ub=support[2]))
        elif support[0] == "ToSimplex":
            chain.append(dict(kind="ToSimplex", D_attr=support[1]))
    return chain


def chain_num_params(chain, D):
    n = 0
    for b in chain:
        n += bijector_num_params(b, D)
    return n


def bijector_num_params(b, D):
    if b["kind"] == "RealNVP":
        return coupling_num_params(D, b["L"], b["U"], b["upper"])
    if b["kind"] == "Affine":
        return 2 * D
    return 0


def fresh_bn_state(chain, D):
    """mean 0 / alpha 1 per BatchNorm (torch_nf/bijectors.py:345-346)."""
    return [(torch.zeros(D), torch.ones(D)) if b["kind"] == "BatchNorm" else None for b in chain]


def base_log_density_f64(omega):
    """float64 numpy base density exactly as the reference forms it: log of the
    product of pdfs (torch_nf/density_estimator.py:369-372)."""
    return np.log(np.prod(np.exp((-np.square(omega)) / 2.0) / np.sqrt(2.0 * np.pi), axis=2))


def normflow_forward(chain, D, params, omega, freeze_bn=False, bn_state=None):
    """Sample direction with injected float64 ``omega`` (M,N,D) in place of the
    reference's in-function ``np.random.normal`` draw.
    torch_nf/density_estimator.py:364-388.  Returns (z f32, log_q_z f64,
    bn_state)."""
    omega = np.asarray(omega, dtype=np.float64)
    z = torch.tensor(omega).float()
    log_q_z = torch.tensor(base_log_density_f64(omega))
    bn_state = list(bn_state) if bn_state is not None else fresh_bn_state(chain, D)
    idx = 0
    for i, b in enumerate(chain):
        k = b["kind"]
        if k == "BatchNorm":
            lm, la = bn_state[i]
            z, log_det, mean, alpha = batchnorm_forward(z, b["eps"], freeze_bn, lm, la)
            bn_state[i] = (mean, alpha)
        elif k == "RealNVP":
            n = bijector_num_params(b, D)
            z, log_det = coupling_forward(z, params[:, idx:idx + n], D, b["L"], b["U"], b["upper"])
            idx += n
        elif k == "Affine":
            z, log_det = affine_forward(z, params[:, idx:idx + 2 * D], D)
            idx += 2 * D
        elif k == "ToInterval":
            z, log_det = tointerval_forward(z, b["lb"], b["ub"])
        elif k == "ToSimplex":
            z, log_det = tosimplex_forward(z, b["D_attr"])
        log_q_z = log_q_z - log_det
    return z, log_q_z, bn_state


def normflow_inverse_and_log_det(chain, D, z, params, bn_state):
    """torch_nf/density_estimator.py:390-406 (params sliced from the end)."""
    idx = chain_num_params(chain, D)
    sum_log_det = torch.zeros((z.shape[0], z.shape[1]), dtype=z.dtype)
    for i in range(len(chain) - 1, -1, -1):
        b = chain[i]
        k = b["kind"]
        if k == "BatchNorm":
            lm, la = bn_state[i]
            z, log_det = batchnorm_inverse(z, lm, la)
        elif k == "RealNVP":
            n = bijector_num_params(b, D)
            z, log_det = coupling_inverse(z, params[:, idx - n:idx], D, b["L"], b["U"], b["upper"])
            idx -= n
        elif k == "Affine":
            z, log_det = affine_inverse(z, params[:, idx - 2 * D:idx], D)
            idx -= 2 * D
        elif k == "ToInterval":
            z, log_det = tointerval_inverse(z, b["lb"], b["ub"])
        else:
            raise TypeError("bijector %s has no inverse" % k)
        sum_log_det = sum_log_det + log_det
    return z, sum_log_det


def normflow_log_prob(chain, D, z, params, bn_state):
    """torch_nf/density_estimator.py:408-416."""
    z0, sum_log_det = normflow_inverse_and_log_det(chain, D, z, params, bn_state)
    log_q_z = torch.sum(-(z0 ** 2), dim=2) / 2.0 - D * math.log(math.sqrt(2.0 * math.pi))
    return log_q_z - sum_log_det


# ---------------------------------------------------------------------------
# Hyper-network of ConditionalDensityEstimator
#                         (torch_nf/conditional_density_estimator.py:19-40,94)
# ---------------------------------------------------------------------------
def param_net(x, weights):
    """``weights`` = [(W (out,in), b (out,)), ...] in nn.Linear convention;
    Tanh after every layer but the last."""
    h = x
    for i, (W, b) in enumerate(weights):
        h = F.linear(h, W, b)
        if i + 1 < len(weights):
            h = torch.tanh(h)
    return h


# ---------------------------------------------------------------------------
# bf16-conditioner emulation (checker for the tcgen05 path; not reference code)
# ---------------------------------------------------------------------------
def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def coupling_bf16_emulated(z, params, D, num_layers, num_units, transform_upper=True, inverse=False):
    """Same layer as coupling_forward / coupling_inverse (bijectors.py:145-206) with
    the conditioner's GEMM operands rounded to bf16 (weights, the conditioning
    half and every tanh output) and fp32 accumulation, fp32 biases, fp32 affine
    transform and log-det: what the tensor-core kernel computes, up to the
    hardware tanh approximation and summation order."""
    d_in, d_out = coupling_dims(D, transform_upper)
    z1, z2 = _split(z, D, transform_upper)
    M = params.shape[0]
    U = num_units
    sizes = [(d_in, U, True)] + [(U, U, True)] * (num_layers - 1) + [(U, d_out, False)]
    t = s = _bf16(z1)
    off = 0
    for (K, J, act) in sizes:
        Wt = _bf16(params[:, off:off + K * J].reshape(M, K, J)); off += K * J
        Ws = _bf16(params[:, off:off + K * J].reshape(M, K, J)); off += K * J
        bt = params[:, off:off + J].reshape(M, 1, J); off += J
        bs = params[:, off:off + J].reshape(M, 1, J); off += J
        t = torch.matmul(t, Wt) + bt
        s = torch.matmul(s, Ws) + bs
        if act:
            t = _bf16(torch.tanh(t))
            s = _bf16(torch.tanh(s))
    z2 = (z2 - t) / torch.exp(s) if inverse else t + z2 * torch.exp(s)
    return _join(z1, z2, transform_upper), torch.sum(s, dim=2)


# ---------------------------------------------------------------------------
# MAF                                    (torch_nf/bijectors.py:597-806)
# ---------------------------------------------------------------------------
def maf_num_params(D, num_layers, num_units):
    """torch_nf/bijectors.py:804-806."""
    return 2 * (2 * D * num_units + (num_layers - 1) * num_units ** 2)


def _maf_nets(z, params, masks, D, num_layers, U):
    """Masked, bias-free shift / log-scale nets (bijectors.py:698-740, 766-796).
    ``masks``: list of (1, K, J) 0/1 tensors, one per layer (shared by both nets)."""
    M = params.shape[0]
    sizes = [(D, U)] + [(U, U)] * (num_layers - 1) + [(U, D)]
    mu = al = z
    off = 0
    for i, (K, J) in enumerate(sizes):
        Wm = masks[i].to(params.dtype) * params[:, off:off + K * J].reshape(M, K, J); off += K * J
        Wa = masks[i].to(params.dtype) * params[:, off:off + K * J].reshape(M, K, J); off += K * J
        mu = torch.matmul(mu, Wm)
        al = torch.matmul(al, Wa)
        if i + 1 < len(sizes):
            mu = torch.tanh(mu)
            al = torch.tanh(al)
    return mu, al


def maf_forward(z, params, masks, D, num_layers, num_units):
    """D-1 passes z <- u*exp(alpha(z)) + mu(z); log-det of the last pass (bijectors.py:742-756)."""
    u = z
    al = None
    for _ in range(D - 1):
        mu, al = _maf_nets(z, params, masks, D, num_layers, num_units)
        z = u * torch.exp(al) + mu
    return z, torch.sum(al, dim=2)


def maf_inverse(z, params, masks, D, num_layers, num_units):
    """bijectors.py:758-764."""
    mu, al = _maf_nets(z, params, masks, D, num_layers, num_units)
    return (z - mu) / torch.exp(al), torch.sum(al, dim=2)
