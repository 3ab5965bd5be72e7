"""Hyper-network fusion (tnf_cde_logprob): ConditionalDensityEstimator.log_prob with the last Linear of param_net
evaluated inside the flow kernel, against the unfused path (params materialised, per-bijector kernels) and the
CPU oracle fed with the reference-style params = param_net(x) (conditional_density_estimator.py:101-104)."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as O
import torch_nf_b200.density_estimator as de
from torch_nf_b200 import _lib, config
from torch_nf_b200.bijectors import ToInterval
from torch_nf_b200.conditional_density_estimator import ConditionalDensityEstimator

pytestmark = pytest.mark.gpu


def _cde(D, D_x, hidden, support, seed=0):
    torch.manual_seed(seed)
    np.random.seed(seed)
    sup = ToInterval(D, [-2.0] * D, [2.0] * D) if support else None
    nf = de.NormFlow(D, True, "coupling", 1, 2, 15, sup)
    cde = ConditionalDensityEstimator(nf, D_x, hidden)
    with torch.no_grad():      # torch's default Linear init gives flow parameters of std ~0.3: far from an identity flow,
        cde.param_net[-1].bias.normal_(0.0, 0.2)      # and still conditioned well enough for a 1e-4 comparison in fp32
    return nf, cde


def _fused_and_unfused(cde, z, x):
    with torch.no_grad():
        config.set_cde_fusion(True)
        n0 = _lib.launch_count()
        a = cde.log_prob(z, x)
        n_fused = _lib.launch_count() - n0
        config.set_cde_fusion(False)
        try:
            n0 = _lib.launch_count()
            b = cde.log_prob(z, x)
            n_unfused = _lib.launch_count() - n0
        finally:
            config.set_cde_fusion(True)
    return a, b, n_fused, n_unfused


@pytest.fixture(params=["tc", "cc"], autouse=True)
def _variant(request):
    """Every test runs with both producers of the parameter rows: tcgen05 (default) and CUDA-core FMA."""
    config.set_cde_variant(request.param)
    yield request.param
    config.set_cde_variant("tc")


@pytest.mark.parametrize("D,D_x,hidden,support,M", [(6, 2, [64, 64], True, 1000), (8, 8, [100], False, 4096 + 77),
                                                    (2, 3, [16], False, 5), (4, 1, [7, 33], True, 129), (6, 2, [64, 64], True, 1),
                                                    (6, 2, [64, 64], True, 148 * 128 * 3 + 17), (8, 8, [100], False, 148 * 128 + 1)])
def test_fused_logprob_matches_unfused_and_oracle(D, D_x, hidden, support, M):
    nf, cde = _cde(D, D_x, hidden, support)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(M, D_x, generator=g)
    z = (torch.rand(M, 1, D, generator=g) * 3.6 - 1.8) if support else torch.randn(M, 1, D, generator=g)
    # BatchNorm statistics as a sampling pass leaves them (otherwise identity: mean 0, alpha 1)
    with torch.no_grad():
        cde(x, N=1)
    lp_f, lp_u, n_fused, n_unfused = _fused_and_unfused(cde, z, x)
    assert lp_f.shape == lp_u.shape == (M, 1)
    assert n_fused <= 2 and n_unfused > n_fused, (n_fused, n_unfused)    # (pack on first use +) ONE kernel
    err = ((lp_f - lp_u).abs() / lp_u.abs().clamp(min=1)).max().item()
    assert err <= config.FP32_TOL_LOGP, err
    # oracle: the reference's own data flow, params materialised on the CPU
    lb, ub = -2.0 * np.ones(D), 2.0 * np.ones(D)
    chain = O.build_chain(D, "coupling", 1, 2, 15, ("ToInterval", lb, ub) if support else None)
    with torch.no_grad():
        params = cde.param_net(x)
    st = []
    for b in nf.bijectors:
        st.append((b.get_last_mean().float().cpu(), b.get_last_alpha().float().cpu()) if b.name == "BatchNorm" else None)
    lp_o = O.normflow_log_prob(chain, D, z, params, st)
    err_o = ((lp_f.cpu() - lp_o).abs() / lp_o.abs().clamp(min=1)).max().item()
    assert err_o <= config.FP32_TOL_LOGP, err_o


def test_fused_path_tracks_weight_updates_and_falls_back():
    """The packed last layer is rebuilt after an in-place parameter update; calls the kernel cannot serve
    (autograd, several samples per context) take the unfused path."""
    nf, cde = _cde(6, 2, [64, 64], True)
    g = torch.Generator().manual_seed(2)
    M = 300
    x = torch.randn(M, 2, generator=g)
    z = torch.rand(M, 1, 6, generator=g) * 3.6 - 1.8
    lp0, _, _, _ = _fused_and_unfused(cde, z, x)
    with torch.no_grad():
        cde.param_net[-1].weight.add_(0.05)
        cde.param_net[0].bias.add_(0.1)
    lp1, lp1_u, _, _ = _fused_and_unfused(cde, z, x)
    assert (lp1 - lp0).abs().max().item() > 1e-3
    assert ((lp1 - lp1_u).abs() / lp1_u.abs().clamp(min=1)).max().item() <= 2e-5
    # autograd: unfused, differentiable
    lp_g = cde.log_prob(z, x)
    assert lp_g.requires_grad
    (-lp_g.mean()).backward()
    assert cde.param_net[-1].weight.grad is not None
    # N > 1 samples per context: unfused
    with torch.no_grad():
        z2 = torch.rand(M, 3, 6, generator=g) * 3.6 - 1.8
        n0 = _lib.launch_count()
        lp2 = cde.log_prob(z2, x)
        assert lp2.shape == (M, 3) and _lib.launch_count() - n0 > 2
