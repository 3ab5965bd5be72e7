"""Deterministic synthetic weights and inputs for parity tests and bench.py.

The default ``xavier_normal_`` init of a ``(1, D_params)`` leaf
(reference torch_nf/density_estimator.py:352-356) has std ~1e-3 at 1.3 M
parameters: a near-identity flow that would hide conditioner errors.  The
generator below scales each conditioner matrix by its fan-in
so activations stay O(1) through the chain.  It uses the legacy
``numpy.random.RandomState`` stream (bit-stable across numpy releases), so a
seed fully identifies a weight set and golden fixtures only store seeds.
"""
import numpy as np


def coupling_dims(D, transform_upper):
    h = D // 2
    d_in, d_out = h, h
    if D % 2 == 1:
        if transform_upper:
            d_out += 1
        else:
            d_in += 1
    return d_in, d_out


def chain_spec(bijectors):
    """[(kind, L, U, upper)] from objects exposing the reference attributes
    (``name``, ``num_layers``, ``num_units``, ``transform_upper``)."""
    spec = []
    for b in bijectors:
        if b.name == "RealNVP":
            spec.append(("RealNVP", b.num_layers, b.num_units, b.transform_upper))
        else:
            spec.append((b.name, 0, 0, False))
    return spec


def synthetic_params(spec, D, M=1, seed=0, g_hidden=1.0, g_last=0.3, bias_std=0.1,
                     affine_std=0.1, dtype=np.float32):
    """Flat ``(M, D_params)`` parameter array in the reference's packing order
    (chain order; per RealNVP layer W_t, W_s, b_t, b_s; Affine alpha, shift)."""
    rs = np.random.RandomState(seed)
    cols = []
    for (kind, L, U, upper) in spec:
        if kind == "RealNVP":
            d_in, d_out = coupling_dims(D, upper)
            sizes = [(d_in, U, g_hidden)] + [(U, U, g_hidden)] * (L - 1) + [(U, d_out, g_last)]
            for (K, J, g) in sizes:
                std = g / np.sqrt(K)
                cols.append(rs.standard_normal((M, K * J)) * std)       # W_t
                cols.append(rs.standard_normal((M, K * J)) * std)       # W_s
                cols.append(rs.standard_normal((M, J)) * bias_std)      # b_t
                cols.append(rs.standard_normal((M, J)) * bias_std)      # b_s
        elif kind == "Affine":
            cols.append(rs.standard_normal((M, D)) * affine_std)        # alpha
            cols.append(rs.standard_normal((M, D)) * affine_std)        # shift
    if not cols:
        return np.zeros((M, 0), dtype=dtype)
    return np.concatenate(cols, axis=1).astype(dtype)


def synthetic_noise(M, N, D, seed=1):
    """float64 base noise omega ~ N(0,1), the injected stand-in for the
    reference's ``np.random.normal`` draw (density_estimator.py:366)."""
    return np.random.RandomState(seed).standard_normal((M, N, D))
