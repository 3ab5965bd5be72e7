"""ctypes binding of the C-ABI library (include/tnf.h).

The library is the product: there is no Python/PyTorch fallback for any entry
point.  Loading fails loudly when ``_C.so`` is missing and cannot be built.
"""
import ctypes
import os
import re

from . import _build

c_void_p, c_int, c_int64, c_double, c_size_t, c_uint64 = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_size_t, ctypes.c_uint64)

TNF_F32, TNF_F64 = 0, 1
TNF_FORWARD, TNF_INVERSE = 0, 1
TNF_LD_WRITE, TNF_LD_ADD, TNF_LD_SUB = 0, 1, -1
TNF_TC_BF16, TNF_TC_FP32 = 0, 1

P, I, L, Dbl, Z, U64 = c_void_p, c_int, c_int64, c_double, c_size_t, c_uint64

# name -> (restype, argtypes); mirrors include/tnf.h declaration by declaration
PROTOTYPES = {
    "tnf_abi_version": (I, []),
    "tnf_last_error": (ctypes.c_char_p, []),
    "tnf_launch_count": (L, []),
    "tnf_coupling": (I, [P, P, P, P, L, L, L, I, I, I, I, I, I, I, P]),
    "tnf_coupling_bwd": (I, [P, P, L, P, P, P, P, L, L, L, I, I, I, I, I, I, P]),
    "tnf_coupling_bwd_overwrite": (I, [P, P, L, P, P, P, P, L, L, L, I, I, I, I, I, I, P]),
    "tnf_maf": (I, [P, P, P, P, L, P, L, L, I, I, I, I, I, I, P]),
    "tnf_maf_bwd": (I, [P, P, L, P, P, P, P, P, L, L, L, I, I, I, I, I, P]),
    "tnf_tc_supported": (I, [I, I, I, I]),
    "tnf_tc_selftest_gemm": (I, [P, P, P, I, I, I, P]),
    "tnf_tc_packed_bytes": (Z, [I, I, I, I]),
    "tnf_tc_pack": (I, [P, P, I, I, I, I, I, P]),
    "tnf_coupling_tc": (I, [P, P, P, P, L, I, I, I, I, I, I, P, P, P, P, I, I, P, P]),
    "tnf_tc_bwd_supported": (I, [I, I, I]),
    "tnf_tc_bwd_packed_bytes": (Z, [I, I, I]),
    "tnf_tc_bwd_workspace_bytes": (Z, [L, I, I, I]),
    "tnf_tc_bwd_pack": (I, [P, P, I, I, I, I, P]),
    "tnf_coupling_tc_bwd": (I, [P, P, P, P, P, P, L, I, I, I, I, I, P, P, P]),
    "tnf_affine": (I, [P, P, P, P, L, L, L, I, I, I, P]),
    "tnf_affine_bwd": (I, [P, P, L, P, P, P, P, L, L, L, I, I, I, P]),
    "tnf_colstats_workspace_bytes": (Z, [I]),
    "tnf_colstats": (I, [P, L, I, P, P, I, P]),
    "tnf_bn_finalize": (I, [P, I, Dbl, P, P, P, I, P]),
    "tnf_bn_apply": (I, [P, P, P, P, L, I, I, I, P]),
    "tnf_bn_bwd_sums": (I, [P, P, L, I, P, P, I, P]),
    "tnf_bn_bwd_apply": (I, [P, P, P, P, P, P, P, L, I, I, P]),
    "tnf_fold_colaffine": (I, [P, P, I, P, P, P, P, P, I, P]),
    "tnf_colaffine": (I, [P, P, P, P, L, I, P]),
    "tnf_tointerval": (I, [P, P, P, P, L, I, I, I, I, P]),
    "tnf_tointerval_bwd": (I, [P, P, P, P, P, L, I, I, I, P]),
    "tnf_tosimplex": (I, [P, P, P, L, I, I, I, I, P]),
    "tnf_tosimplex_bwd": (I, [P, P, P, P, L, I, I, I, P]),
    "tnf_accum_bcast": (I, [P, P, L, L, I, P]),
    "tnf_base_logprob": (I, [P, P, P, L, P, L, I, I, P]),
    "tnf_base_logprob_bwd": (I, [P, P, P, L, I, I, P]),
    "tnf_base_sample": (I, [P, P, L, I, U64, U64, P]),
    "tnf_base_logq": (I, [P, P, L, I, P]),
    "tnf_finish_logq": (I, [P, P, P, L, L, I, P]),
}



class Bijector(ctypes.Structure):
    """tnf_bijector_t (include/tnf.h)."""
    _fields_ = [("kind", c_int), ("num_layers", c_int), ("num_units", c_int), ("transform_upper", c_int),
                ("param_offset", c_int64), ("packed", c_void_p), ("bn_mean", c_void_p), ("bn_alpha", c_void_p),
                ("bn_log_det", c_void_p), ("bn_eps", c_double), ("consts", c_void_p), ("ev_start", c_void_p),
                ("ev_stop", c_void_p)]


TNF_BIJ_REALNVP, TNF_BIJ_BATCHNORM, TNF_BIJ_AFFINE, TNF_BIJ_TOINTERVAL, TNF_BIJ_TOSIMPLEX = 0, 1, 2, 3, 4
ALLREDUCE_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_int, c_void_p)
TNF_PEER_MAX, TNF_PEER_SLOT = 8, 520


class Peer(ctypes.Structure):
    """tnf_peer_t (include/tnf.h): BatchNorm statistics exchanged over NVLink peer memory."""
    _fields_ = [("rank", c_int), ("world", c_int), ("stats", c_void_p * TNF_PEER_MAX), ("flags", c_void_p * TNF_PEER_MAX),
                ("seq", ctypes.c_ulonglong)]

PROTOTYPES.update({
    "tnf_chain_workspace_bytes": (Z, [L, L, I]),
    "tnf_chain_logprob": (I, [P, I, P, P, L, L, L, I, I, P, P, Z, P]),
    "tnf_chain_sample": (I, [P, I, P, L, L, L, I, I, P, U64, U64, I, ALLREDUCE_FN, P, P, P, P, P, P, Z, P]),
    "tnf_cde_supported": (I, [P, I, I, I]),
    "tnf_cde_packed_bytes": (Z, [L, I, I]),
    "tnf_cde_pack": (I, [P, I, I, P, P, I, P, I, P]),
    "tnf_cde_logprob": (I, [P, I, I, P, I, P, P, L, P, I, P]),
})

_LIB = None


def header_symbols():
    """Function names declared in include/tnf.h (used by the symbol-export test)."""
    hdr = os.path.join(os.path.dirname(_build.HERE), "include", "tnf.h")
    with open(hdr) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(tnf_[a-z0-9_]+)\s*\(", text)))


def lib():
    """Load (building first if the sources changed and nvcc is present)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if not _build.is_current():
        if _build._nvcc() is not None:
            path = _build.build()
        elif not os.path.exists(path):
            raise RuntimeError(
                "torch_nf_b200: CUDA library %s is missing and nvcc is not available to build it; "
                "there is no CPU fallback" % path)
    handle = ctypes.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(handle, name)   # AttributeError = missing symbol: fail loudly
        fn.restype = res
        fn.argtypes = args
    if handle.tnf_abi_version() != 3:
        raise RuntimeError("torch_nf_b200: ABI version mismatch")
    _LIB = handle
    return _LIB


def check(rc, what):
    if rc == 0:
        return
    msg = lib().tnf_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError("%s: %s (code %d)" % (what, msg, rc))
    raise RuntimeError("%s: CUDA error %d: %s" % (what, rc, msg))


def launch_count():
    return int(lib().tnf_launch_count())
