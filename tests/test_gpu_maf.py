"""GPU parity tests of the MAF bijector and the 'AR' NormFlow against the
reference's golden vectors and the oracle.  Follows tests/test_bijectors.py:125-200."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as O
import torch_nf_b200.density_estimator as de
from torch_nf_b200.bijectors import MAF

pytestmark = pytest.mark.gpu
T = torch.tensor


def close(a, b, rtol, atol):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("nm", ["d4_f64", "d20_f64", "d6_f32", "d5_a_f32"])
def test_MAF_golden(golden, nm):
    g = golden("maf")
    D, L, U, M, N, seed = [int(v) for v in g[nm + "_cfg"]]
    np.random.seed(seed)
    maf = MAF(D, L, U)
    params, z_in = T(g[nm + "_params"]), T(g[nm + "_z_in"])
    tol = 1e-10 if params.dtype == torch.float64 else 2e-5
    z, ld = maf(z_in, params)
    close(z, g[nm + "_z_fwd"], tol, tol); close(ld, g[nm + "_ld_fwd"], tol, tol)
    z, ld = maf.inverse_and_log_det(z_in, params)
    close(z, g[nm + "_z_inv"], tol, tol); close(ld, g[nm + "_ld_inv"], tol, tol)


def test_MAF_reference_style():
    """Round trip D=4 and D=20, float64, SSE < 1e-6 (tests/test_bijectors.py:172-199)."""
    for D in (4, 20):
        np.random.seed(0)
        maf = MAF(D, 2, 20)
        M, N = 10, 5
        params = T(np.random.normal(0.0, 0.1, (M, maf.count_num_params())))
        z_in = T(np.random.normal(0.0, 1.0, (M, N, D)))
        z, log_det = maf(z_in, params)
        assert z.shape == (M, N, D) and log_det.shape == (M, N)
        z_inv, log_det_inv = maf.inverse_and_log_det(z, params)
        assert np.sum((z_in.numpy() - z_inv.numpy()) ** 2) < 1e-6
        assert np.sum((log_det.numpy() - log_det_inv.numpy()) ** 2) < 1e-6


def test_AR_flow_golden(golden):
    g = golden("flow_ar")
    D, L, U, M, N, seed = [int(v) for v in g["cfg"]]
    np.random.seed(seed)
    nf = de.NormFlow(D, True, "AR", 1, L, U)
    params = T(g["params"])
    assert nf.D_params == params.shape[1]
    z, lq = nf.forward(params, N, omega=g["omega"])
    close(z, g["z"], 2e-5, 2e-5); close(lq, g["log_q_z"], 1e-4, 1e-4)
    close(nf.bijectors[1].get_last_mean(), g["bn_mean"][0], 1e-4, 2e-5)
    lp = nf.log_prob(T(g["z"]), params)
    close(lp, g["log_prob"], 1e-4, 1e-4)
    # default-architecture unconditional flow: sample / log_prob self-consistency (test_density_estimators.py:232-237)
    nf = de.NormFlow(4, False, "AR", num_layers=2, num_units=20)
    z, log_q_z = nf(10)
    lp = nf.log_prob(z)
    assert np.sum(np.square(log_q_z.detach().numpy() - lp.detach().numpy())) < 1e-2


def test_MAF_inverse_backward():
    g = torch.Generator().manual_seed(5)
    D, L, U, M, N = 5, 2, 12, 4, 6
    np.random.seed(3)
    maf = MAF(D, L, U)
    masks = maf.Ms
    for dtype, tol in ((torch.float64, 1e-9), (torch.float32, 2e-4)):
        params = (torch.randn(M, maf.count_num_params(), generator=g, dtype=torch.float64) * 0.3).to(dtype)
        z = torch.randn(M, N, D, generator=g, dtype=torch.float64).to(dtype)
        wz = torch.randn(M, N, D, generator=g, dtype=torch.float64).to(dtype)
        wl = torch.randn(M, N, generator=g, dtype=torch.float64).to(dtype)
        grads = []
        for fn in (maf.inverse_and_log_det, lambda zz, pp: O.maf_inverse(zz, pp, masks, D, L, U)):
            zz, pp = z.clone().requires_grad_(True), params.clone().requires_grad_(True)
            y, ld = fn(zz, pp)
            ((y * wz).sum() + (ld * wl).sum()).backward()
            grads.append((zz.grad, pp.grad))
        for a, b in zip(grads[0], grads[1]):
            np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=tol * 10, atol=tol)
    # training through the AR flow's log_prob
    nf = de.NormFlow(D, True, "AR", 1, L, U)
    p = (torch.randn(8, nf.D_params, generator=g) * 0.2).requires_grad_(True)
    loss = -nf.log_prob(torch.randn(8, 1, D, generator=g), p).mean()
    loss.backward()
    assert torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0
