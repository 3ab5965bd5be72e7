"""Tensor-level wrappers over the C-ABI kernels (no autograd here).

Every function takes CUDA tensors, validates shape / dtype / contiguity in
Python (the C side only checks pointers and ranges), enqueues on torch's
current stream and returns torch tensors it allocated.  torch is used for
device memory and streams only; all arithmetic happens in ``_C.so``.
"""
import torch

from . import _lib
from ._lib import TNF_F32, TNF_F64, TNF_FORWARD, TNF_INVERSE, TNF_LD_WRITE, TNF_LD_ADD, TNF_LD_SUB  # noqa: F401

_DT = {torch.float32: TNF_F32, torch.float64: TNF_F64}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "torch_nf_b200 computes on a CUDA device (sm_100a) only; no CUDA device is visible "
            "and there is no CPU fallback")


def to_device(t, dtype=None):
    """Move a tensor to the current CUDA device (the reference's own tests hand
    CPU tensors to bijectors; they are staged to HBM, never computed on CPU)."""
    require_cuda()
    if not t.is_cuda:
        t = t.cuda()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


def to_host(t):
    """Device -> host through pinned memory (PyTorch's caching host allocator
    recycles the pinned blocks, so steady-state calls do not re-pin)."""
    if not t.is_cuda:
        return t
    if t.requires_grad:            # keep the autograd edge for differentiable results
        return t.cpu()
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out


def to_like(t, device):
    if t.device == device:
        return t
    if device.type == "cpu":
        return to_host(t)
    return t.to(device)


# bench.py hook: when set to a list, (start, end) CUDA events are recorded around
# every tensor-core coupling launch on the launching stream.
kernel_timer = None
bwd_kernel_timer = None     # the same for the launches of the tensor-core backward kernel (tnf_coupling_tc_bwd)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError("torch_nf_b200 kernels take float32 or float64 tensors, got %s" % t.dtype)


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _check3(z):
    if z.dim() != 3:
        raise ValueError("z must have shape (M, N, D), got %s" % (tuple(z.shape),))
    return z if z.is_contiguous() else z.contiguous()


def param_view(params, n_needed):
    """(tensor, row_stride) for a parameter slice whose rows may be strided."""
    if params.dim() != 2:
        raise ValueError("params must have shape (M, >=|theta|), got %s" % (tuple(params.shape),))
    if params.shape[1] < n_needed:
        raise ValueError("params has %d columns, bijector needs %d" % (params.shape[1], n_needed))
    if params.shape[1] > 1 and params.stride(1) != 1:
        params = params.contiguous()
    stride = params.stride(0) if params.shape[0] > 1 else 0
    return params, stride


def _match_rows(z, params):
    """Reference broadcasting: params with one row serve every m of z."""
    M = z.shape[0]
    Mp = params.shape[0]
    if Mp != M and Mp != 1:
        raise ValueError("params has %d rows but z has M=%d" % (Mp, M))
    return Mp


# ----------------------------------------------------------------- coupling
def coupling_num_params(D, U, L, upper):
    """Parameters one RealNVP layer consumes (reference bijectors.py:244-262)."""
    h = D // 2
    d_in, d_out = (h, D - h) if upper else (D - h, h)
    return 2 * (d_in * U + d_out * U + d_out + U + (L - 1) * (U + 1) * U)


def coupling(z, params, D, U, L, upper, direction, ld=None, accum=TNF_LD_WRITE):
    """RealNVP layer on the exact CUDA-core path. Returns (z_out, log_det (M,N))."""
    z = _check3(z)
    M, N, _ = z.shape
    params, pstride = param_view(params, coupling_num_params(D, U, L, upper))
    Mp = _match_rows(z, params)
    z_out = torch.empty_like(z)
    if ld is None:
        ld = torch.empty((M, N), dtype=z.dtype, device=z.device)
        accum = TNF_LD_WRITE
    Mk, Nk = (1, M * N) if Mp == 1 else (M, N)
    rc = _lib.lib().tnf_coupling(z.data_ptr(), z_out.data_ptr(), ld.data_ptr(), params.data_ptr(), pstride, Mk, Nk,
                                 D, U, L, int(upper), direction, accum, _dt(z), _stream())
    _lib.check(rc, "tnf_coupling")
    return z_out, ld


def coupling_bwd(z_in, params, g_z_out, g_ld, g_params, D, U, L, upper, direction, overwrite=False):
    """Accumulates into ``g_params`` (same row layout as ``params``); returns g_z_in.  ``overwrite``: the slice is
    WRITTEN instead (conditional regime: one row per m, few samples per row; needs no zero-filled buffer); when the
    shape does not allow it the slice is zeroed here and accumulated into."""
    z_in = _check3(z_in)
    M, N, _ = z_in.shape
    params, pstride = param_view(params, coupling_num_params(D, U, L, upper))
    Mp = _match_rows(z_in, params)
    assert g_params.shape[0] == Mp and (g_params.shape[1] <= 1 or g_params.stride(1) == 1)
    gstride = g_params.stride(0) if Mp > 1 else 0
    g_z_out = None if g_z_out is None else g_z_out.contiguous()
    g_ld = None if g_ld is None else g_ld.contiguous()
    g_z = torch.empty_like(z_in)
    Mk, Nk = (1, M * N) if Mp == 1 else (M, N)
    args = (z_in.data_ptr(), params.data_ptr(), pstride, _ptr(g_z_out), _ptr(g_ld), g_z.data_ptr(), g_params.data_ptr(),
            gstride, Mk, Nk, D, U, L, int(upper), direction, _dt(z_in), _stream())
    if overwrite:
        if Mp > 1 and Nk <= 32 and _lib.lib().tnf_coupling_bwd_overwrite(*args) == 0:
            return g_z
        g_params.zero_()
    rc = _lib.lib().tnf_coupling_bwd(*args)
    _lib.check(rc, "tnf_coupling_bwd")
    return g_z


# ----------------------------------------------------------------- MAF
def maf(z, params, mask, D, U, L, direction, ld=None, accum=TNF_LD_WRITE):
    z = _check3(z)
    M, N, _ = z.shape
    params, pstride = param_view(params, mask.numel())
    Mp = _match_rows(z, params)
    z_out = torch.empty_like(z)
    if ld is None:
        ld = torch.empty((M, N), dtype=z.dtype, device=z.device)
        accum = TNF_LD_WRITE
    Mk, Nk = (1, M * N) if Mp == 1 else (M, N)
    rc = _lib.lib().tnf_maf(z.data_ptr(), z_out.data_ptr(), ld.data_ptr(), params.data_ptr(), pstride,
                            mask.data_ptr(), Mk, Nk, D, U, L, direction, accum, _dt(z), _stream())
    _lib.check(rc, "tnf_maf")
    return z_out, ld


def maf_bwd(z_in, params, mask, g_z_out, g_ld, g_params, D, U, L, direction):
    z_in = _check3(z_in)
    M, N, _ = z_in.shape
    params, pstride = param_view(params, mask.numel())
    Mp = _match_rows(z_in, params)
    gstride = g_params.stride(0) if Mp > 1 else 0
    g_z_out = None if g_z_out is None else g_z_out.contiguous()
    g_ld = None if g_ld is None else g_ld.contiguous()
    g_z = torch.empty_like(z_in)
    Mk, Nk = (1, M * N) if Mp == 1 else (M, N)
    rc = _lib.lib().tnf_maf_bwd(z_in.data_ptr(), params.data_ptr(), pstride, mask.data_ptr(), _ptr(g_z_out),
                                _ptr(g_ld), g_z.data_ptr(), g_params.data_ptr(), gstride, Mk, Nk, D, U, L,
                                direction, _dt(z_in), _stream())
    _lib.check(rc, "tnf_maf_bwd")
    return g_z


# ----------------------------------------------------------------- tensor-core coupling
TC_PRECISION = {"bf16": _lib.TNF_TC_BF16, "fp32_tc": _lib.TNF_TC_FP32}


def tc_supported(D, U, L, precision="bf16"):
    return bool(_lib.lib().tnf_tc_supported(D, U, L, TC_PRECISION[precision]))


def tc_pack(params_row, D, U, L, upper, precision="bf16"):
    """fp32 parameter row -> packed UMMA operand images (bf16, or fp16 hi/lo for the fp32-parity mode)."""
    nbytes = _lib.lib().tnf_tc_packed_bytes(D, U, L, TC_PRECISION[precision])
    if nbytes == 0:
        raise ValueError("tensor-core path: D=%d U=%d L=%d precision=%s not supported" % (D, U, L, precision))
    packed = torch.empty(nbytes, dtype=torch.uint8, device=params_row.device)
    p = params_row.reshape(-1)
    need = coupling_num_params(D, U, L, upper)
    if p.numel() < need:
        raise ValueError("tc_pack: parameter row has %d values, the layer needs %d" % (p.numel(), need))
    if p.dtype != torch.float32 or not p.is_contiguous():
        p = p.float().contiguous()
    rc = _lib.lib().tnf_tc_pack(p.data_ptr(), packed.data_ptr(), D, U, L, int(upper), TC_PRECISION[precision],
                                _stream())
    _lib.check(rc, "tnf_tc_pack")
    return packed


def coupling_tc(z, packed, D, U, L, upper, direction, ld=None, accum=TNF_LD_WRITE, pre_scale=None, pre_shift=None,
                want_stats=False, out=None, precision="bf16", variant=0, debug=None):
    """Returns (z_out, log_det) or, with ``want_stats`` (D <= 128), (z_out, log_det, sums) where ``sums`` is the
    float64 [sum | sumsq | rows] buffer of the OUTPUT columns (the next BatchNorm's statistics)."""
    z2 = z.reshape(-1, D)
    if z2.dtype != torch.float32 or not z2.is_contiguous():
        raise TypeError("tensor-core coupling takes contiguous float32 z")
    rows = z2.shape[0]
    z_out = torch.empty_like(z2) if out is None else out
    if ld is None:
        ld = torch.empty(rows, dtype=torch.float32, device=z.device)
        accum = TNF_LD_WRITE
    sums = ws = None
    if want_stats:
        sums = torch.empty(2 * D + 1, dtype=torch.float64, device=z.device)
        ws = _workspace(D, z.device)
    timer = kernel_timer
    if timer is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = _lib.lib().tnf_coupling_tc(z2.data_ptr(), z_out.data_ptr(), ld.data_ptr(), packed.data_ptr(), rows, D, U, L,
                                    int(upper), direction, accum, _ptr(pre_scale), _ptr(pre_shift), _ptr(sums),
                                    _ptr(ws), TC_PRECISION[precision], int(variant), _ptr(debug), _stream())
    if timer is not None:
        e1.record()
        timer.append((e0, e1))
    _lib.check(rc, "tnf_coupling_tc")
    if want_stats:
        return z_out.view(z.shape), ld, sums
    return z_out.view(z.shape), ld


# ----------------------------------------------------------------- tensor-core coupling backward
def tc_bwd_supported(D, U, L):
    return bool(_lib.lib().tnf_tc_bwd_supported(D, U, L))


def tc_bwd_pack(params_row, D, U, L, upper):
    """fp32 parameter row -> forward + transposed bf16 operand images of the backward kernel."""
    nbytes = _lib.lib().tnf_tc_bwd_packed_bytes(D, U, L)
    if nbytes == 0:
        raise ValueError("tensor-core backward: D=%d U=%d L=%d not supported" % (D, U, L))
    p = params_row.reshape(-1)
    need = coupling_num_params(D, U, L, upper)
    if p.numel() < need:
        raise ValueError("tc_bwd_pack: parameter row has %d values, the layer needs %d" % (p.numel(), need))
    if p.dtype != torch.float32 or not p.is_contiguous():
        p = p.float().contiguous()
    packed = torch.empty(nbytes, dtype=torch.uint8, device=p.device)
    rc = _lib.lib().tnf_tc_bwd_pack(p.data_ptr(), packed.data_ptr(), D, U, L, int(upper), _stream())
    _lib.check(rc, "tnf_tc_bwd_pack")
    return packed


TC_BWD_MAX_ROWS = 1 << 20      # rows per kernel call: bounds the bf16 workspace (4.2 KB per row at D=64, U=256)


def _mm_t_f32(a, b):
    """a^T @ b: a (rows, m), b (rows, n) bf16 -> (m, n) float32 (a library GEMM over the batch dimension K = rows)."""
    try:
        return torch.mm(a.t(), b, out_dtype=torch.float32)
    except (TypeError, RuntimeError):
        return torch.mm(a.t(), b).float()


def coupling_tc_bwd(z_in, packed_bwd, g_z_out, g_ld, g_params_row, D, U, L, upper, direction, pre_scale=None, pre_shift=None):
    """Backward of the shared-weight coupling layer on tensor cores (tnf_coupling_tc_bwd).  Returns g_z_in and
    ACCUMULATES the parameter gradient into ``g_params_row`` (flat, the layer's slice, layout of bijectors.py:224-235).
    The data gradient and the recompute run in the CUDA kernel; the weight gradients are GEMMs over the batch
    (K = rows) on the bf16 matrices the kernel leaves in its workspace: per net and layer ONE GEMM, the activation
    operand carrying a ones column so that the bias gradient is a row of the product.
    ``pre_scale`` / ``pre_shift`` (D floats): the layer acted on ``z_in * pre_scale + pre_shift`` (a folded BatchNorm);
    the returned gradient is w.r.t. ``z_in``."""
    z2 = z_in.reshape(-1, D)
    if z2.dtype != torch.float32 or not z2.is_contiguous():
        raise TypeError("tensor-core coupling backward takes contiguous float32 z")
    rows_all = z2.shape[0]
    gz2 = None if g_z_out is None else g_z_out.reshape(-1, D).contiguous()
    gl = None if g_ld is None else g_ld.reshape(-1).contiguous()
    g_z = torch.empty_like(z2)
    DH = D // 2
    UP = U + 16
    gp = g_params_row.reshape(-1)
    # offsets of the layer's pieces in the flat row: per layer [W_t | W_s | b_t | b_s]
    o0, o1 = 0, 2 * DH * U + 2 * U
    o2 = o1 + 2 * U * U + 2 * U
    lay = [(o0, DH, U), (o1, U, U), (o2, U, DH)]
    for lo in range(0, rows_all, TC_BWD_MAX_ROWS):
        hi = min(rows_all, lo + TC_BWD_MAX_ROWS)
        rows = hi - lo
        nbytes = _lib.lib().tnf_tc_bwd_workspace_bytes(rows, D, U, L)
        ws = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=z2.device)
        btimer = bwd_kernel_timer
        if btimer is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = _lib.lib().tnf_coupling_tc_bwd(z2[lo:hi].data_ptr(), packed_bwd.data_ptr(),
                                            0 if gz2 is None else gz2[lo:hi].data_ptr(),
                                            0 if gl is None else gl[lo:hi].data_ptr(), g_z[lo:hi].data_ptr(),
                                            ws.data_ptr(), rows, D, U, L, int(upper), direction, _ptr(pre_scale),
                                            _ptr(pre_shift), _stream())
        if btimer is not None:
            e1.record()
            btimer.append((e0, e1, rows))
        _lib.check(rc, "tnf_coupling_tc_bwd")
        mats = ws[:8 * rows * UP].view(2, 4, rows, UP)       # [net][h1, h2, d1, d2]
        d3 = ws[8 * rows * UP: 8 * rows * UP + 2 * rows * DH].view(2, rows, DH)
        # xa = (x | 1 | 0..) as the layer saw it, written by the kernel; padded to 64 / 128 columns: the library's
        # 40-column product takes 0.49 ms, the 64-column one 0.28 (profiles/r02_lines/r02z_wgrad_gemm_probe2.json)
        XW = 64 if DH + 1 <= 64 else 128
        xa = ws[8 * rows * UP + 2 * rows * DH:].view(rows, XW)
        # per net and layer ONE GEMM: (activation | 1 | 0..)^T delta; rows 0..K-1 of the product are dW, row K is db.
        # (measured, profiles/scripts/wgrad_gemm_probe.py: the full-pitch activation matrix with its ones column costs
        # no more than the plain one and saves a column-sum pass over delta; a batched GEMM over both nets is 3x slower)
        for net in range(2):
            acts = (xa, mats[net, 0], mats[net, 1])
            deltas = (mats[net, 2][:, :U], mats[net, 3][:, :U], d3[net])
            for l, (off, K, J) in enumerate(lay):
                g = _mm_t_f32(acts[l], deltas[l])
                gp[off + net * K * J: off + (net + 1) * K * J] += g[:K].reshape(-1)
                boff = off + 2 * K * J + net * J
                gp[boff: boff + J] += g[K]
        del ws
    return g_z.view(z_in.shape)


# ----------------------------------------------------------------- affine
def affine(z, params, D, direction, want_ld=True):
    z = _check3(z)
    M, N, _ = z.shape
    params, pstride = param_view(params, 2 * D)
    Mp = _match_rows(z, params)
    z_out = torch.empty_like(z)
    ld = torch.empty((Mp, 1), dtype=z.dtype, device=z.device) if want_ld else None
    Mk, Nk = (1, M * N) if Mp == 1 else (M, N)
    rc = _lib.lib().tnf_affine(z.data_ptr(), z_out.data_ptr(), _ptr(ld), params.data_ptr(), pstride, Mk, Nk, D,
                               direction, _dt(z), _stream())
    _lib.check(rc, "tnf_affine")
    return z_out, ld


def affine_bwd(z_in, params, g_z_out, g_ld, g_params, D, direction):
    z_in = _check3(z_in)
    M, N, _ = z_in.shape
    params, pstride = param_view(params, 2 * D)
    Mp = _match_rows(z_in, params)
    gstride = g_params.stride(0) if Mp > 1 else 0
    g_z_out = None if g_z_out is None else g_z_out.contiguous()
    g_ld = None if g_ld is None else g_ld.contiguous()
    g_z = torch.empty_like(z_in)
    Mk, Nk = (1, M * N) if Mp == 1 else (M, N)
    rc = _lib.lib().tnf_affine_bwd(z_in.data_ptr(), params.data_ptr(), pstride, _ptr(g_z_out), _ptr(g_ld),
                                   g_z.data_ptr(), g_params.data_ptr(), gstride, Mk, Nk, D, direction, _dt(z_in),
                                   _stream())
    _lib.check(rc, "tnf_affine_bwd")
    return g_z


# ----------------------------------------------------------------- batch norm
_ws_cache = {}


def _workspace(D, device):
    """Per-CTA partial sums of the column-statistics kernels.  One buffer per (D, device, STREAM): launches on
    different streams must not share partials; launches on one stream are ordered."""
    key = (D, device, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None:
        ws = torch.empty(_lib.lib().tnf_colstats_workspace_bytes(D), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def colstats(z, D):
    """[sum (D) | sum of squares (D) | rows] of the flattened (rows, D) batch, float64."""
    z2 = z.reshape(-1, D)
    z2 = z2 if z2.is_contiguous() else z2.contiguous()
    sums = torch.empty(2 * D + 1, dtype=torch.float64, device=z.device)
    rc = _lib.lib().tnf_colstats(z2.data_ptr(), z2.shape[0], D, sums.data_ptr(), _workspace(D, z.device).data_ptr(),
                                 _dt(z2), _stream())
    _lib.check(rc, "tnf_colstats")
    return sums


def bn_finalize(sums, D, eps, dtype):
    mean = torch.empty(D, dtype=dtype, device=sums.device)
    alpha = torch.empty(D, dtype=dtype, device=sums.device)
    ld = torch.empty((), dtype=dtype, device=sums.device)
    rc = _lib.lib().tnf_bn_finalize(sums.data_ptr(), D, float(eps), mean.data_ptr(), alpha.data_ptr(),
                                    ld.data_ptr(), _DT[dtype], _stream())
    _lib.check(rc, "tnf_bn_finalize")
    return mean, alpha, ld


def bn_apply(z, mean, alpha, D, direction):
    zc = z if z.is_contiguous() else z.contiguous()
    z_out = torch.empty_like(zc)
    rows = zc.numel() // D
    rc = _lib.lib().tnf_bn_apply(zc.data_ptr(), z_out.data_ptr(), mean.data_ptr(), alpha.data_ptr(), rows, D,
                                 direction, _dt(zc), _stream())
    _lib.check(rc, "tnf_bn_apply")
    return z_out


def bn_bwd_sums(g_y, y, D):
    y2 = y.reshape(-1, D)
    y2 = y2 if y2.is_contiguous() else y2.contiguous()
    g2 = g_y.reshape(-1, D)
    g2 = g2 if g2.is_contiguous() else g2.contiguous()
    gs = torch.empty(2 * D + 1, dtype=torch.float64, device=y.device)
    rc = _lib.lib().tnf_bn_bwd_sums(g2.data_ptr(), y2.data_ptr(), y2.shape[0], D, gs.data_ptr(),
                                    _workspace(D, y.device).data_ptr(), _dt(y2), _stream())
    _lib.check(rc, "tnf_bn_bwd_sums")
    return gs


def bn_bwd_apply(g_y, y, alpha, gsums, g_ld, count, D):
    """``count``: 1-element float64 device tensor holding the global row count."""
    yc = y if y.is_contiguous() else y.contiguous()
    gc = None if g_y is None else (g_y if g_y.is_contiguous() else g_y.contiguous())
    g_z = torch.empty_like(yc)
    rows = yc.numel() // D
    rc = _lib.lib().tnf_bn_bwd_apply(_ptr(gc), yc.data_ptr(), alpha.data_ptr(), gsums.data_ptr(), _ptr(g_ld),
                                     count.data_ptr(), g_z.data_ptr(), rows, D, _dt(yc), _stream())
    _lib.check(rc, "tnf_bn_bwd_apply")
    return g_z


# ----------------------------------------------------------------- folded per-column affine maps
FOLD_BN_FWD, FOLD_BN_INV, FOLD_AFF_FWD, FOLD_AFF_INV = 0, 1, 2, 3


def fold_colaffine(pend, kind, a, b, D, ld_accum=None):
    """Compose the pending map ``pend`` = (scale, shift) or None with one more BatchNorm / Affine.
    For Affine kinds ``ld_accum[0] += sum(alpha)`` when given."""
    ps = torch.empty(D, dtype=torch.float32, device=a.device)
    pb = torch.empty(D, dtype=torch.float32, device=a.device)
    rc = _lib.lib().tnf_fold_colaffine(_ptr(pend[0]) if pend else 0, _ptr(pend[1]) if pend else 0, kind,
                                       a.data_ptr(), b.data_ptr(), ps.data_ptr(), pb.data_ptr(), _ptr(ld_accum), D,
                                       _stream())
    _lib.check(rc, "tnf_fold_colaffine")
    return ps, pb


def colaffine(z, pend, D):
    zc = z if z.is_contiguous() else z.contiguous()
    out = torch.empty_like(zc)
    rc = _lib.lib().tnf_colaffine(zc.data_ptr(), out.data_ptr(), pend[0].data_ptr(), pend[1].data_ptr(),
                                  zc.numel() // D, D, _stream())
    _lib.check(rc, "tnf_colaffine")
    return out


# ----------------------------------------------------------------- support layers
def tointerval(z, consts, D, direction, ld=None, accum=TNF_LD_WRITE):
    z = _check3(z)
    M, N, _ = z.shape
    z_out = torch.empty_like(z)
    if ld is None:
        ld = torch.empty((M, N), dtype=z.dtype, device=z.device)
        accum = TNF_LD_WRITE
    rc = _lib.lib().tnf_tointerval(z.data_ptr(), z_out.data_ptr(), ld.data_ptr(), consts.data_ptr(), M * N, D,
                                   direction, accum, _dt(z), _stream())
    _lib.check(rc, "tnf_tointerval")
    return z_out, ld


def tointerval_bwd(z_in, consts, g_z_out, g_ld, D, direction):
    z_in = _check3(z_in)
    M, N, _ = z_in.shape
    g_z_out = None if g_z_out is None else g_z_out.contiguous()
    g_ld = None if g_ld is None else g_ld.contiguous()
    g_z = torch.empty_like(z_in)
    rc = _lib.lib().tnf_tointerval_bwd(z_in.data_ptr(), consts.data_ptr(), _ptr(g_z_out), _ptr(g_ld), g_z.data_ptr(),
                                       M * N, D, direction, _dt(z_in), _stream())
    _lib.check(rc, "tnf_tointerval_bwd")
    return g_z


def tosimplex(z, D_attr, ld=None, accum=TNF_LD_WRITE):
    z = _check3(z)
    M, N, Din = z.shape
    z_out = torch.empty((M, N, Din + 1), dtype=z.dtype, device=z.device)
    if ld is None:
        ld = torch.empty((M, N), dtype=z.dtype, device=z.device)
        accum = TNF_LD_WRITE
    rc = _lib.lib().tnf_tosimplex(z.data_ptr(), z_out.data_ptr(), ld.data_ptr(), M * N, Din, D_attr, accum, _dt(z),
                                  _stream())
    _lib.check(rc, "tnf_tosimplex")
    return z_out, ld


def tosimplex_bwd(z_in, g_z_out, g_ld, D_attr):
    z_in = _check3(z_in)
    M, N, Din = z_in.shape
    g_z_out = None if g_z_out is None else g_z_out.contiguous()
    g_ld = None if g_ld is None else g_ld.contiguous()
    g_z = torch.empty_like(z_in)
    rc = _lib.lib().tnf_tosimplex_bwd(z_in.data_ptr(), _ptr(g_z_out), _ptr(g_ld), g_z.data_ptr(), M * N, Din, D_attr,
                                      _dt(z_in), _stream())
    _lib.check(rc, "tnf_tosimplex_bwd")
    return g_z


# ----------------------------------------------------------------- base density
def accum_bcast(dst, src, div):
    """dst[i] += src[i // div] (flattened), in place."""
    rc = _lib.lib().tnf_accum_bcast(dst.data_ptr(), src.data_ptr(), dst.numel(), int(div), _dt(dst), _stream())
    _lib.check(rc, "tnf_accum_bcast")
    return dst


def base_logprob(z, sub=None, scal=None, scal_div=1):
    z = _check3(z)
    M, N, D = z.shape
    out = torch.empty((M, N), dtype=z.dtype, device=z.device)
    rc = _lib.lib().tnf_base_logprob(z.data_ptr(), _ptr(sub), _ptr(scal), int(scal_div), out.data_ptr(), M * N, D,
                                     _dt(z), _stream())
    _lib.check(rc, "tnf_base_logprob")
    return out


def base_logprob_bwd(z, g_out):
    z = _check3(z)
    M, N, D = z.shape
    g_out = g_out.contiguous()
    g_z = torch.empty_like(z)
    rc = _lib.lib().tnf_base_logprob_bwd(z.data_ptr(), g_out.data_ptr(), g_z.data_ptr(), M * N, D, _dt(z), _stream())
    _lib.check(rc, "tnf_base_logprob_bwd")
    return g_z


def base_sample(M, N, D, seed, offset, device):
    z = torch.empty((M, N, D), dtype=torch.float32, device=device)
    log_q = torch.empty((M, N), dtype=torch.float64, device=device)
    rc = _lib.lib().tnf_base_sample(z.data_ptr(), log_q.data_ptr(), M * N, D, int(seed) & (2 ** 64 - 1),
                                    int(offset) & (2 ** 64 - 1), _stream())
    _lib.check(rc, "tnf_base_sample")
    return z, log_q


def base_logq(omega):
    omega = _check3(omega)
    M, N, D = omega.shape
    log_q = torch.empty((M, N), dtype=torch.float64, device=omega.device)
    rc = _lib.lib().tnf_base_logq(omega.data_ptr(), log_q.data_ptr(), M * N, D, _stream())
    _lib.check(rc, "tnf_base_logq")
    return log_q


def finish_logq(log_q, ld_acc, scal=None, scal_div=1):
    rows = log_q.numel()
    ref = ld_acc if ld_acc is not None else scal
    rc = _lib.lib().tnf_finish_logq(log_q.data_ptr(), _ptr(ld_acc), _ptr(scal), int(scal_div), rows,
                                    _dt(ref) if ref is not None else TNF_F32, _stream())
    _lib.check(rc, "tnf_finish_logq")
    return log_q
