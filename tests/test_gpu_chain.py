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
        # sampling with LIVE BatchNorm statistics: base density + ONE cooperative launch (grid barrier per BatchNorm)
        ((z, lq), (z_p, lq_p), n_chain, n_plan) = _both(lambda: nf.forward(params, N, omega=omega))
        assert n_chain == 2 and n_plan > 4, (n_chain, n_plan)
        assert ((z - z_p).abs() / z_p.abs().clamp(min=1)).max().item() <= 1e-5
        assert (lq - lq_p).abs().max().item() <= 2e-5 * max(1.0, float(lq_p.abs().max()))
        bn = [b for b in nf.bijectors if b.name == "BatchNorm"]
        m_plan = [b.get_last_mean().clone() for b in bn]
        z, lq = nf.forward(params, N, omega=omega)           # (the chain call again: it sets the BatchNorm statistics)
        for b, mp in zip(bn, m_plan):
            assert (b.get_last_mean() - mp).abs().max().item() <= 1e-6
        (lp_c, lp_p, n_chain, n_plan) = _both(lambda: nf.log_prob(z, params))
        assert n_chain == 1 and n_plan > 4, (n_chain, n_plan)
        assert (lp_c - lp_p).abs().max().item() <= 2e-5 * max(1.0, float(lp_p.abs().max()))
        ((zf_c, lq_c), (zf_p, lq_p), n_chain, n_plan) = _both(lambda: nf.forward(params, N, freeze_bn=True, omega=omega))
        assert n_chain == 2 and n_plan > 4, (n_chain, n_plan)       # base density + the fused chain
        assert ((zf_c - zf_p).abs() / zf_p.abs().clamp(min=1)).max().item() <= 1e-5
        assert (lq_c - lq_p).abs().max().item() <= 2e-5 * max(1.0, float(lq_p.abs().max()))
    chain = O.build_chain(D, "coupling", stages, L, U)
    zo, lqo, st = O.normflow_forward(chain, D, params.cpu(), omega)
    assert ((z.cpu() - zo).abs() / zo.abs().clamp(min=1)).max().item() <= 1e-5
    assert ((lq.cpu() - lqo).abs() / lqo.abs().clamp(min=1)).max().item() <= 1e-4
    lpo = O.normflow_log_prob(chain, D, z.cpu(), params.cpu(), st)
    assert ((lp_c.cpu() - lpo).abs() / lpo.abs().clamp(min=1)).max().item() <= 1e-4


@pytest.mark.parametrize("precision,D", [("fp32", 64), ("bf16", 64), ("bf16", 128), ("bf16", 256), ("fp32", 256)])
def test_chain_calls_match_plan_c3(precision, D):
    """Tensor-core chains: in the sample direction the C-side executor's fused fold launches reproduce the host-side
    plan bit for bit; log_prob folds every BatchNorm / Affine with ONE launch and takes the base density inside the last
    executed coupling layer (no z0 store, no base-density pass), so it agrees with the plan to fp32 rounding."""
    stages, L, U, N = 2, 2, 256, 1000
    nf = de.NormFlow(D, True, "coupling", stages, L, U)
    params = T(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=4)).cuda()
    omega = np.random.RandomState(5).standard_normal((1, N, D))
    tnf.set_conditioner_precision(precision)
    try:
        with torch.no_grad():
            ((z_c, lq_c), (z_p, lq_p), n_chain, n_plan) = _both(lambda: nf.forward(params, N, omega=omega))
            assert torch.equal(z_c, z_p) and torch.equal(lq_c, lq_p)
            assert n_chain < n_plan
            (lp_c, lp_p, n_chain, n_plan) = _both(lambda: nf.log_prob(z_c, params))
            err = ((lp_c - lp_p).abs() / lp_p.abs().clamp(min=1)).max().item()
            print("fused base density vs plan: rel %.3g, launches %d vs %d" % (err, n_chain, n_plan))
            assert err <= 2e-6
            # one fold launch + one launch per coupling layer (+ the base density where the layer's kernel cannot fuse it)
            assert n_chain == 2 * stages + 1 + (1 if (precision, D) == ("bf16", 128) else 0), n_chain
    finally:
        tnf.set_conditioner_precision("fp32")


@pytest.mark.parametrize("precision,D,U", [("bf16", 64, 256), ("fp32", 64, 256), ("bf16", 256, 256), ("bf16", 128, 128), ("fp32", 128, 128)])
def test_chain_logprob_ragged_rows(precision, D, U):
    """Fused base density at row counts that leave partial tiles, single tiles, idle CTAs, and several tiles per CTA
    (the per-tile hand-off buffers between the I/O and epilogue warps wrap around)."""
    tnf.set_conditioner_precision(precision)
    config.set_tc_min_rows(1)
    try:
        nf = de.NormFlow(D, True, "coupling", 1, 2, U)
        params = T(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=8)).cuda()
        for N in (1, 127, 129, 300, 148 * 128 + 5, 148 * 128 * 5 + 77):
            omega = np.random.RandomState(9).standard_normal((1, N, D))
            with torch.no_grad():
                z, _ = nf.forward(params, N, omega=omega)
                (lp_c, lp_p, _, _) = _both(lambda: nf.log_prob(z, params))
            assert lp_c.shape == lp_p.shape == (1, N)
            err = ((lp_c - lp_p).abs() / lp_p.abs().clamp(min=1)).max().item()
            assert err <= 2e-6, (N, err)
    finally:
        config.set_tc_min_rows(128)
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
