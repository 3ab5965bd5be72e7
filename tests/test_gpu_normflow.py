"""GPU parity tests of the NormFlow chain (sample, log_prob, frozen BatchNorm,
conditional hyper-network) against the golden vectors of the unmodified
reference and against the CPU oracle.  Structure follows the reference's
tests/test_density_estimators.py and tests/test_conditional_density_estimators.py."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as O
import torch_nf_b200.bijectors as bij
import torch_nf_b200.density_estimator as de
from torch_nf_b200.conditional_density_estimator import ConditionalDensityEstimator
from torch_nf_b200.synthetic import chain_spec, synthetic_params

pytestmark = pytest.mark.gpu
T = torch.tensor


def rel_z(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


def _flow_case(golden, name, support=None, has_logprob=True, device="cpu"):
    g = golden(name)
    D, stages, L, U, M, N, pseed, oseed = [int(v) for v in g["cfg"]]
    nf = de.NormFlow(D, True, "coupling", stages, L, U, support)
    if "params" in g.files:
        params, omega = T(g["params"]), g["omega"]
    else:
        params = T(synthetic_params(chain_spec(nf.bijectors), D, M, seed=pseed))
        np.random.seed(oseed)
        omega = np.random.normal(0.0, 1.0, (M, N, D))
    params = params.to(device)
    assert nf.D_params == params.shape[1]
    # log_prob before any forward sees identity BatchNorm
    z, log_q_z = nf.forward(params, N, omega=omega)
    assert z.dtype == torch.float32 and log_q_z.dtype == torch.float64
    assert z.device == params.device and tuple(z.shape) == g["z"].shape
    zc, lq = z.cpu().numpy(), log_q_z.cpu().numpy()
    assert rel_z(zc, g["z"]) <= 1e-5
    assert rel_z(lq, g["log_q_z"]) <= 1e-4
    bns = [b for b in nf.bijectors if b.name == "BatchNorm"]
    for i, b in enumerate(bns):
        np.testing.assert_allclose(b.get_last_mean().cpu().numpy(), g["bn_mean"][i], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(b.get_last_alpha().cpu().numpy(), g["bn_alpha"][i], rtol=1e-4, atol=1e-6)
    if has_logprob:
        lp = nf.log_prob(T(g["z"]).to(device), params)
        assert lp.dtype == torch.float32
        assert rel_z(lp.cpu().numpy(), g["log_prob"]) <= 1e-4
        # self-consistency exactly as the reference tests it (SSE < 1e-2)
        lp_own = nf.log_prob(z, params)
        assert np.sum(np.square(lq - lp_own.cpu().numpy())) < 1e-2 * max(1.0, M * N / 10.0)
    np.random.seed(oseed + 1)
    omega2 = np.random.normal(0.0, 1.0, (M, N, D))
    z_f, lq_f = nf.forward(params, N, freeze_bn=True, omega=omega2)
    assert rel_z(z_f.cpu().numpy(), g["z_frozen"]) <= 1e-5
    assert rel_z(lq_f.cpu().numpy(), g["log_q_z_frozen"]) <= 1e-4


def test_flow_c1(golden):
    _flow_case(golden, "flow_c1")
    _flow_case(golden, "flow_c1", device="cuda")


def test_flow_c2(golden):
    _flow_case(golden, "flow_c2a")
    _flow_case(golden, "flow_c2b", device="cuda")


def test_flow_c3_fp32(golden):
    _flow_case(golden, "flow_c3", device="cuda")


def test_flow_c5_fp32(golden):
    _flow_case(golden, "flow_c5", device="cuda")


def test_flow_support_layers(golden):
    lb = -2.0 * np.ones(6); ub = 2.0 * np.ones(6)
    _flow_case(golden, "flow_c4", support=bij.ToInterval(6, lb, ub))
    _flow_case(golden, "flow_simplex", support=bij.ToSimplex(6), has_logprob=False)
    nf = de.NormFlow(5, True, "coupling", 1, 2, 15, bij.ToSimplex(6))
    with pytest.raises(TypeError):
        nf.log_prob(torch.rand(2, 3, 6), torch.zeros(2, nf.D_params))   # ToSimplex has no inverse


def test_NormFlow_reference_style():
    """tests/test_density_estimators.py:206-243: unconditional flows, default init,
    device sampler, forward/log_prob self-consistency."""
    D, N = 4, 10
    for arch, stages in (("coupling", 1), ("coupling", 2), ("affine", 1)):
        nf = de.NormFlow(D, False, arch, stages, 2, 20, None)
        z, log_q_z = nf(N)
        assert z.shape[0] == 1 and z.shape[1] == N and z.shape[2] == D
        assert log_q_z.shape[0] == 1 and log_q_z.shape[1] == N
        log_q_z_inv = nf.log_prob(z)
        assert np.sum(np.square(log_q_z.detach().numpy() - log_q_z_inv.detach().numpy())) < 1e-2


def test_device_sampler_statistics():
    """Philox base sampler: N(0,1) moments, reproducible under np.random.seed,
    float64 base density consistent with the draw."""
    nf = de.NormFlow(8, False, "affine")
    nf.params = torch.zeros(1, nf.D_params)
    np.random.seed(5)
    z, lq = nf(200000)
    np.random.seed(5)
    z2, _ = nf(200000)
    assert torch.equal(z, z2)
    zz = z.numpy().reshape(-1)
    assert abs(zz.mean()) < 5e-3 and abs(zz.std() - 1.0) < 5e-3
    assert abs(np.mean(zz ** 3)) < 2e-2 and abs(np.mean(zz ** 4) - 3.0) < 5e-2
    ref = O.base_log_density_f64(z.numpy().astype(np.float64))
    np.testing.assert_allclose(lq.numpy(), ref, rtol=1e-9, atol=1e-9)


def test_ConditionalDensityEstimator(golden):
    """tests/test_conditional_density_estimators.py:15-69 + the reference's own
    hyper-network weights (golden flow_c2b_net)."""
    D, D_x, M, N = 4, 3, 10, 6
    nf = de.NormFlow(D, True, "coupling", 1, 2, 20)
    cde = ConditionalDensityEstimator(nf, D_x, [16, 16])
    x = torch.randn(M, D_x)
    z, log_q_z = cde(x, N)
    assert z.shape == (M, N, D) and log_q_z.shape == (M, N)
    lp = cde.log_prob(z, x)
    assert np.sum(np.square(log_q_z.detach().numpy() - lp.detach().numpy())) < 1e-2

    g, gn = golden("flow_c2b"), golden("flow_c2b_net")
    nf = de.NormFlow(8, True, "coupling", 1, 2, 15)
    cde = ConditionalDensityEstimator(nf, 8, [100])
    cde.param_net.load_state_dict({"linear1.weight": T(gn["pn_linear1_weight"]), "linear1.bias": T(gn["pn_linear1_bias"]),
                                   "linear2.weight": T(gn["pn_linear2_weight"]), "linear2.bias": T(gn["pn_linear2_bias"])})
    with torch.no_grad():
        params = cde.param_net(T(gn["x"]))
    np.testing.assert_allclose(params.numpy(), g["params"], rtol=1e-5, atol=1e-6)
    with torch.no_grad():
        lp = cde.log_prob(T(g["z"]), T(gn["x"]))       # BatchNorm state is identity here
    chain = O.build_chain(8, "coupling", 1, 2, 15)
    lpo = O.normflow_log_prob(chain, 8, T(g["z"]), T(g["params"]), O.fresh_bn_state(chain, 8))
    assert rel_z(lp.numpy(), lpo.numpy()) <= 1e-4


def test_round_trip_full_size_property():
    """Size-independent property at a large batch: log_prob(sample) equals the
    sampler's own log-density, and inverse(forward) returns the base noise."""
    D, stages, U = 16, 2, 32
    nf = de.NormFlow(D, True, "coupling", stages, 2, U)
    params = T(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=7)).cuda()
    N = 1 << 17
    np.random.seed(0)
    with torch.no_grad():
        z, lq = nf.forward(params, N)
        lp = nf.log_prob(z, params)
        z0, sld = nf.inverse_and_log_det(z, params)
    assert torch.isfinite(z).all() and torch.isfinite(lp).all()
    assert float((lq.float() - lp).abs().max()) < 2e-3
    # z0 must be standard normal again
    m = z0.double().mean().item(); s = z0.double().std().item()
    assert abs(m) < 5e-3 and abs(s - 1.0) < 5e-3


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_log_prob_host_pipeline_is_exact(precision):
    """log_prob of HOST samples runs as overlapped row chunks (copy stream / compute stream); the inverse chain has
    no cross-sample coupling, so the result must equal the one-shot device path bit for bit."""
    import torch_nf_b200 as tnf
    from torch_nf_b200 import config
    D, N = 64, 128 * 37 + 19
    nf = de.NormFlow(D, False, "coupling", 2, 2, 64)
    params = torch.tensor(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=4)).cuda()
    old_rows, old_chunks = config.host_pipeline_min_rows(), config.host_pipeline_chunks()
    tnf.set_conditioner_precision(precision)
    try:
        with torch.no_grad():
            z, _ = nf.forward(params, N)                      # sets the BatchNorm statistics
            lp_dev = nf.log_prob(z, params)
            z_host = z.cpu()
            config.set_host_pipeline(min_rows=1 << 30)
            lp_one = nf.log_prob(z_host, params)              # one-shot host path
            config.set_host_pipeline(min_rows=256, chunks=5)
            lp_pipe = nf.log_prob(z_host, params)             # 5 ragged chunks
            lp_pin = nf.log_prob(z_host.pin_memory(), params)
    finally:
        tnf.set_conditioner_precision("fp32")
        config.set_host_pipeline(min_rows=old_rows, chunks=old_chunks)
    assert not lp_pipe.is_cuda and lp_pipe.shape == (1, N)
    assert torch.equal(lp_one, lp_dev.cpu())
    assert torch.equal(lp_pipe, lp_one)
    assert torch.equal(lp_pin, lp_one)
    assert torch.isfinite(lp_pipe).all()
