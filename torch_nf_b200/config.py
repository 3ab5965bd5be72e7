"""Run-time switches of the hot path.

``conditioner precision``
    ``"fp32"`` (default): conditioner GEMMs on CUDA cores in the input dtype;
    matches the reference within fp32 rounding (max|dz|/max(1,|z|) <= 1e-5,
    |d log_prob| <= 1e-4*max(1,|log_prob|)).
    ``"bf16"``: shared-weight coupling layers run on tcgen05 tensor cores with
    bf16 operands, fp32 accumulation, fp32 affine transform and log-det.
    Stated tolerance: max|dz| <= 5e-2, |d log_prob| <= 2e-3 relative.
"""
import os

_precision = os.environ.get("TNF_CONDITIONER_PRECISION", "fp32")
_tc_min_rows = int(os.environ.get("TNF_TC_MIN_ROWS", "128"))


def set_conditioner_precision(mode):
    global _precision
    if mode not in ("fp32", "bf16"):
        raise ValueError('conditioner precision must be "fp32" or "bf16"')
    _precision = mode


def conditioner_precision():
    return _precision


def set_tc_min_rows(n):
    global _tc_min_rows
    _tc_min_rows = int(n)


def tc_min_rows():
    return _tc_min_rows
