"""tnf_chain_logprob / tnf_chain_sample (one C-ABI call per chain) against the per-bijector plan and the oracle."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as O
import torch_nf_b200 as tnf
import torch_nf_b200.density_estimator as de
from torch_nf_b200 import _lib, config
from torch_nf_b200.bijectors import ToInterval
from torch_nf_b200.synthetic import chain_spec, synthetic_params

pytestmark = pytest.mark.gpu
T = torch.tensor


def _both(fn):
    config.set_chain_abi(True)
    try:
        before = _lib.launch_count()
        a = fn()
        n_chain = _lib.launch_count() - before
    finally:
        config.set_chain_abi(True)
    config.set_chain_abi(False)
    try:
        before = _lib.launch_count()
        b = fn()
        n_plan = _lib.launch_count() - before
    finally:
        config.set_chain_abi(True)
    return a, b, n_chain, n_plan


@pytest.mark.parametrize("D,stages,L,U,N", [(2, 1, 2, 15, 1024), (8, 1, 2, 15, 65536), (8, 3, 3, 20, 333), (5, 2, 1, 64, 77)])
def test_small_chain_is_one_launch(D, stages, L, U, N):
    """C1 / C2a-like shared-weight flows: log_prob is ONE kernel launch (z in registers for the whole chain) and
    agrees with the per-bijector plan and the oracle; sampling with remembered BatchNorm statistics likewise."""
    nf = de.NormFlow(D, True, "coupling", stages, L, U)
    params = T(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=2)).cuda()
    omega = np.random.RandomState(3).standard_normal((1, N, D))
    with torch.no_grad():
        z, lq = nf.forward(params, N, omega=omega)           # sets the BatchNorm statistics
        (lp_c, lp_p, n_chain, n_plan) = _both(lambda: nf.log_prob(z, params))
        assert n_chain == 1 and n_plan > 4, (n_chain, n_plan)
        assert (lp_c - lp_p).abs().max().item() <= 2e-5 * max(1.0, float(lp_p.abs().max()))
        ((zf_c, lq_c), (zf_p, lq_p), n_chain, n_plan) = _both(lambda: nf.forward(params, N, freeze_bn=True, omega=omega))
        assert n_chain == 2 and n_plan > 4, (n_chain, n_plan)       # base density + the fused chain
        assert ((zf_c - zf_p).abs() / zf_p.abs().clamp(min=1)).max().item() <= 1e-5
        assert (lq_c - lq_p).abs().max().item() <= 2e-5 * max(1.0, float(lq_p.abs().max()))
    chain = O.build_chain(D, "coupling", stages, L, U)
    zo, lqo, st = O.normflow_forward(chain, D, params.cpu(), omega)
    lpo = O.normflow_log_prob(chain, D, z.cpu(), params.cpu(), st)
    assert ((lp_c.cpu() - lpo).abs() / lpo.abs().clamp(min=1)).max().item() <= 1e-4


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_chain_calls_match_plan_c3(precision):
    """Tensor-core chains: the C-side executor launches the same kernels as the host-side plan (bit-identical)."""
    D, stages, L, U, N = 64, 2, 2, 256, 1000
    nf = de.NormFlow(D, True, "coupling", stages, L, U)
    params = T(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=4)).cuda()
    omega = np.random.RandomState(5).standard_normal((1, N, D))
    tnf.set_conditioner_precision(precision)
    try:
        with torch.no_grad():
            ((z_c, lq_c), (z_p, lq_p), n_chain, n_plan) = _both(lambda: nf.forward(params, N, omega=omega))
            assert torch.equal(z_c, z_p) and torch.equal(lq_c, lq_p)
            (lp_c, lp_p, _, _) = _both(lambda: nf.log_prob(z_c, params))
            assert torch.equal(lp_c, lp_p)
    finally:
        tnf.set_conditioner_precision("fp32")


def test_chain_with_support_layer_and_per_sample_weights():
    """C4-like: conditional weights (one parameter row per sample, N = 1) and a ToInterval support layer."""
    D, M = 6, 4096
    lb, ub = [-2.0] * D, [2.0] * D
    nf = de.NormFlow(D, True, "coupling", 1, 2, 15, ToInterval(D, lb, ub))
    params = T(synthetic_params(chain_spec(nf.bijectors), D, M, seed=6)).cuda()
    omega = np.random.RandomState(7).standard_normal((M, 1, D))
    with torch.no_grad():
        ((z_c, lq_c), (z_p, lq_p), _, _) = _both(lambda: nf.forward(params, 1, omega=omega))
        assert ((z_c - z_p).abs()).max().item() <= 1e-6 and (lq_c - lq_p).abs().max().item() <= 1e-5
        (lp_c, lp_p, _, _) = _both(lambda: nf.log_prob(z_c, params))
        assert (lp_c - lp_p).abs().max().item() <= 1e-5
    assert float(z_c.min()) >= -2.0 and float(z_c.max()) <= 2.0
