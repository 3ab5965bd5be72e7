"""Tensor-core backward of the shared-weight coupling layer (tnf_coupling_tc_bwd) against torch autograd through the
CPU oracle, and the bf16-mode training path of a C3-like chain built on it.

Stated tolerance (SURVEY 8d, bf16 conditioner): rel-L2 <= 1e-2 on the parameter gradient; the same bound is asserted
for the sample gradient.  The oracle differentiates the fp32 reference arithmetic (oracle/flow_oracle.py)."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as O
import torch_nf_b200.density_estimator as de
from torch_nf_b200 import config, ops
from torch_nf_b200._lib import TNF_FORWARD, TNF_INVERSE
from torch_nf_b200.synthetic import chain_spec, synthetic_params

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-2


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.parametrize("D,U,upper,rows", [(64, 256, True, 1024), (64, 256, False, 1000), (64, 128, True, 300),
                                           (128, 256, False, 640), (128, 128, True, 129)])
@pytest.mark.parametrize("direction", [TNF_INVERSE, TNF_FORWARD])
def test_tc_backward_layer(D, U, upper, rows, direction):
    L = 2
    assert ops.tc_bwd_supported(D, U, L)
    spec = [("RealNVP", L, U, upper)]
    params = torch.tensor(synthetic_params(spec, D, 1, seed=D + U))
    g = torch.Generator().manual_seed(7)
    z = torch.randn(1, rows, D, generator=g)
    wz = torch.randn(1, rows, D, generator=g)
    wl = torch.randn(1, rows, generator=g)
    # oracle autograd (fp32 reference arithmetic)
    zo, po = z.clone().requires_grad_(True), params.clone().requires_grad_(True)
    fn = O.coupling_inverse if direction == TNF_INVERSE else O.coupling_forward
    zz, ld = fn(zo, po, D, L, U, upper)
    ((zz * wz).sum() + (ld * wl).sum()).backward()
    # the kernel
    dev = torch.device("cuda")
    pd = params.to(dev)
    packed = ops.tc_bwd_pack(pd[0], D, U, L, upper)
    g_params = torch.zeros_like(pd)
    g_z = ops.coupling_tc_bwd(z.to(dev).contiguous(), packed, wz.to(dev), wl.to(dev), g_params[0], D, U, L, upper, direction)
    torch.cuda.synchronize()
    rz, rp = _rel(g_z.cpu(), zo.grad), _rel(g_params.cpu(), po.grad)
    print("D=%d U=%d upper=%d rows=%d dir=%d: rel-L2 g_z %.2e g_params %.2e" % (D, U, upper, rows, direction, rz, rp))
    assert torch.isfinite(g_z).all() and torch.isfinite(g_params).all()
    assert rz < GRAD_TOL and rp < GRAD_TOL, (rz, rp)
    # per-piece check of the parameter gradient: every weight / bias block of both nets
    off = 0
    h = D // 2
    for (K, J) in ((h, U), (U, U), (U, h)):
        for n_el in (K * J, K * J, J, J):
            a, b = g_params[0, off:off + n_el].cpu(), po.grad[0, off:off + n_el]
            assert _rel(a, b) < 3 * GRAD_TOL, (K, J, n_el, _rel(a, b))
            off += n_el


def test_tc_backward_null_gradients():
    """g_z_out = NULL and g_log_det = NULL are zeros."""
    D, U, L, rows = 64, 256, 2, 256
    params = torch.tensor(synthetic_params([("RealNVP", L, U, True)], D, 1, seed=3)).cuda()
    z = torch.randn(1, rows, D, device="cuda")
    packed = ops.tc_bwd_pack(params[0], D, U, L, True)
    wl = torch.randn(rows, device="cuda")
    ga, gb = torch.zeros_like(params), torch.zeros_like(params)
    z1 = ops.coupling_tc_bwd(z, packed, None, wl, ga[0], D, U, L, True, TNF_INVERSE)
    z2 = ops.coupling_tc_bwd(z, packed, torch.zeros_like(z), wl, gb[0], D, U, L, True, TNF_INVERSE)
    assert torch.equal(z1, z2) and torch.equal(ga, gb)
    gc = torch.zeros_like(params)
    z3 = ops.coupling_tc_bwd(z, packed, None, None, gc[0], D, U, L, True, TNF_INVERSE)
    assert float(z3.abs().max()) == 0.0 and float(gc.abs().max()) == 0.0


def test_c3_training_gradients_bf16():
    """-mean log_prob of a C3-shaped flow (2 stages here) in the bf16-conditioner mode: forward and backward of every
    coupling layer on tensor cores, against oracle autograd; several tiles per CTA and a ragged last tile."""
    D, stages, U, N = 64, 2, 256, 148 * 128 * 2 + 77
    old = config.conditioner_precision()
    config.set_conditioner_precision("bf16")
    try:
        nf = de.NormFlow(D, False, "coupling", stages, 2, U)
        chain = O.build_chain(D, "coupling", stages, 2, U)
        params0 = torch.tensor(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=0))
        with torch.no_grad():
            z, _ = nf.forward(params0, N, omega=np.random.RandomState(1).standard_normal((1, N, D)))
        st = [(b.get_last_mean().float().cpu(), b.get_last_alpha().float().cpu()) if b.name == "BatchNorm" else None
              for b in nf.bijectors]
        # "data" that is NOT distributed as the model: on the model's own samples the expected score is zero, the true
        # gradient is pure sampling noise and a relative error says nothing (measured there: 2.6e-2)
        z = (z.detach() * 1.25 + 0.1).contiguous()
        p1, z1 = params0.clone().requires_grad_(True), z.clone().requires_grad_(True)
        assert all(nf._tc_train(b, p1, z1) for b in nf.bijectors if b.name == "RealNVP")
        launches0 = de.ops._lib.launch_count()
        nf.params = p1                      # conditioner=False: log_prob uses the flow's own parameter leaf
        loss = -nf.log_prob(z1).mean()
        loss.backward()
        assert de.ops._lib.launch_count() > launches0
        p2, z2 = params0.clone().requires_grad_(True), z.clone().requires_grad_(True)
        loss_o = -O.normflow_log_prob(chain, D, z2, p2, st).mean()
        loss_o.backward()
        rp, rz = _rel(p1.grad, p2.grad), _rel(z1.grad, z2.grad)
        print("C3-like bf16 training: loss %.5f vs %.5f, rel-L2 g_params %.2e g_z %.2e" % (loss.item(), loss_o.item(), rp, rz))
        assert abs(loss.item() - loss_o.item()) < config.BF16_TOL_LOGP * max(1.0, abs(loss_o.item()))
        assert rp < GRAD_TOL and rz < GRAD_TOL, (rp, rz)
    finally:
        config.set_conditioner_precision(old)


def test_sample_direction_gradients_bf16():
    """EFN-style loss through the SAMPLE direction (notebooks/two_network_arch.ipynb:77-83: mean(log_q - eta . T(z))) in
    the bf16-conditioner mode: every coupling layer's forward and backward on tensor cores through the per-bijector
    autograd functions, BatchNorm batch statistics differentiated, against oracle autograd."""
    D, stages, U, N = 64, 1, 256, 4096
    old = config.conditioner_precision()
    config.set_conditioner_precision("bf16")
    try:
        nf = de.NormFlow(D, False, "coupling", stages, 2, U)
        chain = O.build_chain(D, "coupling", stages, 2, U)
        params0 = torch.tensor(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=2))
        omega = np.random.RandomState(3).standard_normal((1, N, D))
        w = torch.tensor(np.random.RandomState(4).standard_normal(D).astype(np.float32))
        p1 = params0.clone().requires_grad_(True)
        z, lq = nf.forward(p1, N, omega=omega)
        loss = (lq.float() - (z * w).sum(dim=2)).mean()
        loss.backward()
        p2 = params0.clone().requires_grad_(True)
        zo, lqo, _ = O.normflow_forward(chain, D, p2, omega)
        loss_o = (lqo.float() - (zo * w).sum(dim=2)).mean()
        loss_o.backward()
        rp = _rel(p1.grad, p2.grad)
        print("sample direction, bf16: loss %.5f vs %.5f, max|dz| %.2e, rel-L2 g_params %.2e"
              % (loss.item(), loss_o.item(), float((z.detach() - zo.detach()).abs().max()), rp))
        assert float((z.detach() - zo.detach()).abs().max()) < config.BF16_TOL_Z
        assert abs(loss.item() - loss_o.item()) < config.BF16_TOL_LOGP * max(1.0, abs(loss_o.item()))
        assert rp < GRAD_TOL, rp
    finally:
        config.set_conditioner_precision(old)


def test_tc_backward_full_size_against_cuda_cores():
    """At more than BASELINE's full batch (2^20 + 2^17 + 5 rows: two workspace chunks, 55 tiles per CTA, a ragged last
    tile) the tensor-core backward against the EXACT CUDA-core backward kernel on the same inputs: the same stated bound."""
    D, U, L, rows = 64, 256, 2, (1 << 20) + (1 << 17) + 5
    params = torch.tensor(synthetic_params([("RealNVP", L, U, True)], D, 1, seed=11)).cuda()
    g = torch.Generator(device="cuda").manual_seed(5)
    z = torch.randn(1, rows, D, device="cuda", generator=g)
    gz = torch.randn(1, rows, D, device="cuda", generator=g)
    gl = torch.randn(1, rows, device="cuda", generator=g)
    gp_tc = torch.zeros_like(params)
    gz_tc = ops.coupling_tc_bwd(z, ops.tc_bwd_pack(params[0], D, U, L, True), gz, gl, gp_tc[0], D, U, L, True, TNF_INVERSE)
    gp_cc = torch.zeros_like(params)
    gz_cc = ops.coupling_bwd(z, params, gz, gl, gp_cc, D, U, L, True, TNF_INVERSE)
    torch.cuda.synchronize()
    rz, rp = _rel(gz_tc, gz_cc), _rel(gp_tc, gp_cc)
    print("full size (%d rows): tensor-core vs CUDA-core backward rel-L2 g_z %.2e g_params %.2e" % (rows, rz, rp))
    assert rz < GRAD_TOL and rp < GRAD_TOL, (rz, rp)
    # linearity in the output gradients (a size-independent property): bwd(2 g) = 2 bwd(g)
    gp2 = torch.zeros_like(params)
    gz2 = ops.coupling_tc_bwd(z, ops.tc_bwd_pack(params[0], D, U, L, True), 2 * gz, 2 * gl, gp2[0], D, U, L, True, TNF_INVERSE)
    assert _rel(gz2, 2 * gz_tc) < 1e-6 and _rel(gp2, 2 * gp_tc) < 2e-3


def test_mixed_precision_training_fp32_forward_bf16_backward():
    """config.set_training_backward("bf16") in the default fp32 mode: the forward keeps fp32 parity (tensor cores, split
    operands), the gradients come from the bf16 tensor-core backward: loss at the fp32 tolerance, gradients at the bf16 one."""
    D, stages, U, N = 64, 2, 256, 148 * 128 + 19
    assert config.conditioner_precision() == "fp32" and not config.tc_backward_enabled()
    config.set_training_backward("bf16")
    try:
        assert config.tc_backward_enabled()
        nf = de.NormFlow(D, False, "coupling", stages, 2, U)
        chain = O.build_chain(D, "coupling", stages, 2, U)
        params0 = torch.tensor(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=6))
        with torch.no_grad():
            z, _ = nf.forward(params0, N, omega=np.random.RandomState(2).standard_normal((1, N, D)))
        st = [(b.get_last_mean().float().cpu(), b.get_last_alpha().float().cpu()) if b.name == "BatchNorm" else None
              for b in nf.bijectors]
        z = (z.detach() * 1.25 + 0.1).contiguous()
        p1 = params0.clone().requires_grad_(True)
        nf.params = p1
        loss = -nf.log_prob(z).mean()
        loss.backward()
        p2 = params0.clone().requires_grad_(True)
        loss_o = -O.normflow_log_prob(chain, D, z, p2, st).mean()
        loss_o.backward()
        rp = _rel(p1.grad, p2.grad)
        print("mixed precision: loss %.6f vs %.6f, rel-L2 g_params %.2e" % (loss.item(), loss_o.item(), rp))
        assert abs(loss.item() - loss_o.item()) < config.FP32_TOL_LOGP * max(1.0, abs(loss_o.item()))
        assert rp < GRAD_TOL, rp
    finally:
        config.set_training_backward("auto")
    assert not config.tc_backward_enabled()
