"""Backward kernels vs torch autograd through the CPU oracle (float64 and float32)."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as O
import torch_nf_b200.bijectors as bij
import torch_nf_b200.density_estimator as de
from torch_nf_b200.conditional_density_estimator import ConditionalDensityEstimator
from torch_nf_b200.synthetic import chain_spec, synthetic_params

pytestmark = pytest.mark.gpu


def _grads(fn_ours, fn_oracle, inputs, wz, wl):
    """d/d(inputs) of sum(z*wz) + sum(log_det*wl) through both implementations."""
    outs = []
    for fn in (fn_ours, fn_oracle):
        xs = [x.clone().requires_grad_(True) for x in inputs]
        z, ld = fn(*xs)
        loss = (z * wz).sum() + (ld * wl.reshape(ld.shape) if ld.numel() == wl.numel() else ld * wl.sum()).sum()
        loss.backward()
        outs.append([x.grad for x in xs])
    return outs


def _check(ours, ref, rtol, atol):
    for a, b in zip(ours, ref):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=rtol, atol=atol)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 2e-4)])
@pytest.mark.parametrize("D,L,U,upper,M,N", [(4, 2, 15, True, 6, 5), (5, 1, 15, False, 3, 7), (8, 3, 17, True, 1, 40),
                                            (6, 2, 15, False, 33, 1)])
def test_coupling_backward(dtype, tol, D, L, U, upper, M, N):
    g = torch.Generator().manual_seed(D * 100 + L)
    b = bij.RealNVP(D, L, U, transform_upper=upper)
    P = b.count_num_params()
    params = (torch.randn(M, P + 2, generator=g, dtype=torch.float64) * 0.3).to(dtype)
    z = torch.randn(M, N, D, generator=g, dtype=torch.float64).to(dtype)
    wz = torch.randn(M, N, D, generator=g, dtype=torch.float64).to(dtype)
    wl = torch.randn(M, N, generator=g, dtype=torch.float64).to(dtype)
    for ours_fn, ref_fn in ((b.forward_and_log_det, O.coupling_forward), (b.inverse_and_log_det, O.coupling_inverse)):
        ours, ref = _grads(ours_fn, lambda zz, pp: ref_fn(zz, pp, D, L, U, upper), [z, params], wz, wl)
        _check(ours, ref, tol * 10, tol)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-4)])
def test_affine_backward(dtype, tol):
    g = torch.Generator().manual_seed(1)
    D = 5
    a = bij.Affine(D)
    for M, N in ((4, 9), (1, 300), (40, 1), (5000, 1), (4100, 3)):      # the last two: thread-per-(m, d) kernel
        params = (torch.randn(M, 2 * D, generator=g, dtype=torch.float64) * 0.5).to(dtype)
        z = torch.randn(M, N, D, generator=g, dtype=torch.float64).to(dtype)
        wz = torch.randn(M, N, D, generator=g, dtype=torch.float64).to(dtype)
        wl = torch.randn(M, 1, generator=g, dtype=torch.float64).to(dtype)
        for ours_fn, ref_fn in ((a.forward_and_log_det, O.affine_forward), (a.inverse_and_log_det, O.affine_inverse)):
            ours, ref = _grads(ours_fn, lambda zz, pp: ref_fn(zz, pp, D), [z, params], wz, wl)
            _check(ours, ref, tol * 10, tol * (10 if N > 100 else 1))


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 2e-4)])
def test_batchnorm_backward(dtype, tol):
    """Gradient flows through the batch statistics (reference bijectors.py:402-417)."""
    g = torch.Generator().manual_seed(2)
    D, M, N = 6, 3, 50
    z = (torch.randn(M, N, D, generator=g, dtype=torch.float64) * 2.0 + 1.0).to(dtype)
    wz = torch.randn(M, N, D, generator=g, dtype=torch.float64).to(dtype)
    wl = torch.randn(1, generator=g, dtype=torch.float64).to(dtype)
    bn = bij.BatchNorm(D)
    ours, ref = _grads(lambda zz: bn(zz), lambda zz: O.batchnorm_forward(zz)[:2], [z], wz, wl)
    _check(ours, ref, tol * 10, tol)
    # stored statistics are constants
    ours, ref = _grads(lambda zz: bn(zz, use_last=True),
                       lambda zz: O.batchnorm_forward(zz, 1e-5, True, bn.get_last_mean().to(dtype), bn.get_last_alpha().to(dtype))[:2],
                       [z], wz, wl * 0)
    _check(ours, ref, tol * 10, tol)
    ours, ref = _grads(lambda zz: bn.inverse_and_log_det(zz),
                       lambda zz: O.batchnorm_inverse(zz, bn.get_last_mean().to(dtype), bn.get_last_alpha().to(dtype)),
                       [z], wz, wl * 0)
    _check(ours, ref, tol * 10, tol)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-8), (torch.float32, 5e-4)])
def test_support_layer_backward(dtype, tol):
    g = torch.Generator().manual_seed(3)
    D, M, N = 6, 4, 11
    lb = np.array([-0.5, -np.inf, -0.5, -np.inf, 1.0, 0.0])
    ub = np.array([0.5, 0.5, np.inf, np.inf, 4.0, np.inf])
    ti = bij.ToInterval(D, lb, ub)
    z = torch.randn(M, N, D, generator=g, dtype=torch.float64).to(dtype)
    wz = torch.randn(M, N, D, generator=g, dtype=torch.float64).to(dtype)
    wl = torch.randn(M, N, generator=g, dtype=torch.float64).to(dtype)
    ours, ref = _grads(lambda zz: ti(zz), lambda zz: O.tointerval_forward(zz, lb, ub), [z], wz, wl)
    _check(ours, ref, tol * 10, tol)
    y = O.tointerval_forward(z, lb, ub)[0].detach()
    ours, ref = _grads(lambda zz: ti.inverse_and_log_det(zz), lambda zz: O.tointerval_inverse(zz, lb, ub), [y], wz, wl)
    _check(ours, ref, tol * 20, tol * 5)
    ts = bij.ToSimplex(D + 1)
    wz2 = torch.randn(M, N, D + 1, generator=g, dtype=torch.float64).to(dtype)
    ours, ref = _grads(lambda zz: ts(zz), lambda zz: O.tosimplex_forward(zz, D + 1), [z], wz2, wl)
    _check(ours, ref, tol * 10, tol)


def test_normflow_training_gradients():
    """d(-mean log_prob)/d params and d(mean(log_q - sum z))/d params for a conditional flow with a
    ToInterval support layer (the C4 training configuration at test size), against oracle autograd."""
    D, M, N = 6, 24, 1
    lb, ub = -2.0 * np.ones(D), 2.0 * np.ones(D)
    nf = de.NormFlow(D, True, "coupling", 1, 2, 15, bij.ToInterval(D, lb, ub))
    chain = O.build_chain(D, "coupling", 1, 2, 15, ("ToInterval", lb, ub))
    rs = np.random.RandomState(0)
    params0 = torch.tensor(synthetic_params(chain_spec(nf.bijectors), D, M, seed=4))
    zobs = torch.tensor(rs.uniform(-1.9, 1.9, (M, N, D)).astype(np.float32))
    # SNPE-style loss (notebooks/LFI_learning_rules.ipynb:295-306)
    p1 = params0.clone().requires_grad_(True)
    loss = -nf.log_prob(zobs, p1).mean()
    loss.backward()
    p2 = params0.clone().requires_grad_(True)
    loss_o = -O.normflow_log_prob(chain, D, zobs, p2, O.fresh_bn_state(chain, D)).mean()
    loss_o.backward()
    assert abs(loss.item() - loss_o.item()) < 1e-4 * max(1.0, abs(loss_o.item()))
    rel = (p1.grad - p2.grad).norm() / p2.grad.norm()
    assert rel < 1e-4, rel
    # EFN-style loss through the sample direction, BatchNorm batch statistics included
    omega = rs.standard_normal((M, 8, D))
    p1 = params0.clone().requires_grad_(True)
    z, lq = nf.forward(p1, 8, omega=omega)
    loss = (lq.float() - z.sum(dim=2)).mean()
    loss.backward()
    p2 = params0.clone().requires_grad_(True)
    zo, lqo, _ = O.normflow_forward(chain, D, p2, omega)
    loss_o = (lqo.float() - zo.sum(dim=2)).mean()
    loss_o.backward()
    assert abs(loss.item() - loss_o.item()) < 1e-4 * max(1.0, abs(loss_o.item()))
    rel = (p1.grad - p2.grad).norm() / p2.grad.norm()
    assert rel < 1e-4, rel


def test_conditional_training_step():
    """One Adam step of the hyper-network on -mean(log_prob) decreases the loss and matches the oracle's
    parameter update (Adam lr 1e-4 as in notebooks/LFI_learning_rules.ipynb:296)."""
    torch.manual_seed(0)
    D, D_x, M = 4, 3, 64
    nf = de.NormFlow(D, True, "coupling", 1, 2, 15)
    cde = ConditionalDensityEstimator(nf, D_x, [16])
    x = torch.randn(M, D_x)
    z = torch.randn(M, 1, D)
    opt = torch.optim.Adam(cde.parameters(), lr=1e-2)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        loss = -cde.log_prob(z, x).mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


@pytest.mark.parametrize("conditioner_rows", [1, 37])
def test_chain_logprob_autograd_node(conditioner_rows):
    """log_prob as ONE autograd node (_ChainLogProbFn): gradients w.r.t. the parameters (shared row and one row per
    context) AND w.r.t. the samples, against oracle autograd; the per-bijector path gives the same numbers."""
    D, N = 6, 5
    M = conditioner_rows if conditioner_rows > 1 else 3
    lb, ub = -2.0 * np.ones(D), 2.0 * np.ones(D)
    nf = de.NormFlow(D, True, "coupling", 2, 2, 15, bij.ToInterval(D, lb, ub))
    chain = O.build_chain(D, "coupling", 2, 2, 15, ("ToInterval", lb, ub))
    rs = np.random.RandomState(1)
    params0 = torch.tensor(synthetic_params(chain_spec(nf.bijectors), D, conditioner_rows, seed=5))
    if conditioner_rows > 1:      # trailing extra columns are legal and ignored (tests/test_bijectors.py:89-93): zero gradient
        params0 = torch.cat([params0, torch.tensor(rs.standard_normal((conditioner_rows, 3)).astype(np.float32))], dim=1)
    z0 = torch.tensor(rs.uniform(-1.8, 1.8, (M, N, D)).astype(np.float32))
    w = torch.tensor(rs.standard_normal((M, N)).astype(np.float32))
    with torch.no_grad():      # non-trivial remembered BatchNorm statistics
        nf.forward(params0, 64)
    st = [(b.get_last_mean().float().cpu(), b.get_last_alpha().float().cpu()) if b.name == "BatchNorm" else None for b in nf.bijectors]
    p1, z1 = params0.clone().requires_grad_(True), z0.clone().requires_grad_(True)
    assert nf._chain_grad_ok()
    (nf.log_prob(z1, p1) * w).sum().backward()
    p2, z2 = params0.clone().requires_grad_(True), z0.clone().requires_grad_(True)
    (O.normflow_log_prob(chain, D, z2, p2, st) * w).sum().backward()
    assert ((p1.grad - p2.grad).norm() / p2.grad.norm()).item() < 1e-4
    assert ((z1.grad - z2.grad).norm() / z2.grad.norm()).item() < 1e-4
    if conditioner_rows > 1:
        assert float(p1.grad[:, nf.D_params:].abs().max()) == 0.0 and torch.isfinite(p1.grad).all()
    # the bijector-by-bijector autograd path
    p3 = params0.clone().requires_grad_(True)
    nf._chain_grad_ok = lambda: False
    (nf.log_prob(z0, p3) * w).sum().backward()
    assert ((p3.grad - p1.grad).norm() / p1.grad.norm()).item() < 1e-5
