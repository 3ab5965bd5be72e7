"""Run-time switches of the hot path.

``conditioner precision``
    ``"fp32"`` (default): the reference's precision.  Shared-weight coupling layers with
    D in {64, 128}, U in {128, 256} run on tcgen05 tensor cores in the fp32-PARITY mode
    (operands split into fp16 hi + lo parts, three MMAs per product, fp32 accumulation,
    tanh / exp to fp32 accuracy: coupling_tc6.cu); every other case (per-sample weights,
    float64, odd or small D, autograd) runs the exact CUDA-core kernel in the input dtype.
    Stated tolerance, both: max|dz|/max(1,|z|) <= FP32_TOL_Z = 1e-5 and
    |d log_prob| <= FP32_TOL_LOGP = 1e-4*max(1,|log_prob|) on the reference's golden vectors.
    ``"fp32_cc"``: the same tolerance with the CUDA-core kernel everywhere (~50x slower at C3).
    ``"bf16"``: shared-weight coupling layers run on tcgen05 tensor cores with
    bf16 operands, fp32 accumulation, fp32 affine transform and log-det.
    Stated tolerance (ONE statement, asserted as written by tests/test_gpu_tc.py, per
    coupling layer and for whole chains up to C3's 8 and C5's 16 layers):
    ``max|dz| <= BF16_TOL_Z = 5e-2`` and
    ``|d log_prob| <= BF16_TOL_LOGP = 2e-3 * max(1, |log_prob|)`` (same for log_q);
    the log-det of a single coupling layer: ``<= BF16_TOL_LD = 5e-2`` absolute.
    Measured against the reference / oracle (profiles/scripts/parity_measure.py): C3 chain
    2.7e-2 and 1.1e-3 over 2^16 rows, C5 chain 4.2e-2 and 7.3e-4 over 2^13 rows.

``host pipeline``
    ``log_prob`` of HOST-resident samples (no autograd, one shared parameter row)
    with at least ``host_pipeline_min_rows()`` rows is cut into row chunks whose
    host->device copies overlap the inverse chain of the previous chunk.  The
    inverse chain has no cross-sample coupling (BatchNorm uses its stored
    statistics), so the result is identical to the one-shot path.
"""
import os

BF16_TOL_Z = 5e-2        # max |dz|, bf16-conditioner mode
BF16_TOL_LOGP = 2e-3     # |d log_prob| / max(1, |log_prob|), bf16-conditioner mode
BF16_TOL_LD = 5e-2       # |d log_det| of ONE coupling layer (sum of D/2 scale outputs), absolute
FP32_TOL_Z = 1e-5        # max |dz| / max(1, |z|), fp32 modes, on the reference's golden vectors
FP32_TOL_LOGP = 1e-4     # |d log_prob| / max(1, |log_prob|), fp32 modes

_precision = os.environ.get("TNF_CONDITIONER_PRECISION", "fp32")
_tc_min_rows = int(os.environ.get("TNF_TC_MIN_ROWS", "128"))


def set_conditioner_precision(mode):
    global _precision
    if mode not in ("fp32", "fp32_cc", "bf16"):
        raise ValueError('conditioner precision must be "fp32", "fp32_cc" or "bf16"')
    _precision = mode


def conditioner_precision():
    return _precision


def tc_precision():
    """Tensor-core kernel family of the current mode ("bf16" / "fp32_tc"), or None (CUDA cores only)."""
    return {"bf16": "bf16", "fp32": "fp32_tc"}.get(_precision)


def set_tc_min_rows(n):
    global _tc_min_rows
    _tc_min_rows = int(n)


def tc_min_rows():
    return _tc_min_rows


_training_backward = os.environ.get("TNF_TRAINING_BACKWARD", "auto")


def set_training_backward(mode):
    """Backward of shared-weight coupling layers on the differentiable path.
    ``"auto"`` (default): the tensor-core backward (tnf_coupling_tc_bwd, bf16 conditioner, gradients within rel-L2 1e-2)
    in the ``"bf16"`` conditioner mode, the exact CUDA-core backward (~100x slower at C3) in the fp32 modes.
    ``"bf16"``: the tensor-core backward in EVERY tensor-core mode - with ``"fp32"`` the forward values keep fp32 parity
    (fp16 hi / lo split on tcgen05) while the gradients carry the bf16 tolerance: mixed-precision training.
    ``"exact"``: the CUDA-core backward everywhere."""
    global _training_backward
    if mode not in ("auto", "bf16", "exact"):
        raise ValueError('training backward must be "auto", "bf16" or "exact"')
    _training_backward = mode


def training_backward():
    return _training_backward


def tc_backward_enabled():
    """The tensor-core backward is in use for the current settings."""
    if _training_backward == "exact":
        return False
    if _training_backward == "bf16":
        return tc_precision() is not None
    return _precision == "bf16"


_host_pipeline_min_rows = int(os.environ.get("TNF_HOST_PIPELINE_MIN_ROWS", str(1 << 17)))
_host_pipeline_chunks = int(os.environ.get("TNF_HOST_PIPELINE_CHUNKS", "4"))   # measured at 2^20 x 64: 2: 16.5, 3: 16.0, 4: 15.8, 8: 16.2, 16: 18.7 ms per e2e step


def set_host_pipeline(min_rows=None, chunks=None):
    global _host_pipeline_min_rows, _host_pipeline_chunks
    if min_rows is not None:
        _host_pipeline_min_rows = int(min_rows)
    if chunks is not None:
        _host_pipeline_chunks = max(1, int(chunks))


def host_pipeline_min_rows():
    return _host_pipeline_min_rows


def host_pipeline_chunks():
    return _host_pipeline_chunks


_chain_abi = os.environ.get("TNF_CHAIN_ABI", "1") != "0"


def set_chain_abi(on):
    """Route no-autograd float32 ``forward`` / ``log_prob`` through tnf_chain_sample / tnf_chain_logprob (default) or
    through the per-bijector host plan (diagnostics)."""
    global _chain_abi
    _chain_abi = bool(on)


def chain_abi():
    return _chain_abi


_cde_fusion = os.environ.get("TNF_CDE_FUSION", "1") != "0"
_cde_variant = os.environ.get("TNF_CDE_VARIANT", "tc")


def set_cde_variant(v):
    """"tc" (default): the fused kernel forms the parameter rows on tcgen05 tensor cores (fp16 hi / lo split, fp32
    parity); "cc": with fp32 FMAs on CUDA cores."""
    global _cde_variant
    if v not in ("tc", "cc"):
        raise ValueError('cde variant must be "tc" or "cc"')
    _cde_variant = v


def cde_variant():
    return _cde_variant


def set_cde_fusion(on):
    """``ConditionalDensityEstimator.log_prob`` without autograd, one sample per context: evaluate the hyper-network's
    last Linear inside the flow kernel (tnf_cde_logprob; the (M, D_params) parameter matrix never reaches HBM) where
    the chain has a compiled shape (default), or always materialise ``params`` as the reference does."""
    global _cde_fusion
    _cde_fusion = bool(on)


def cde_fusion():
    return _cde_fusion
