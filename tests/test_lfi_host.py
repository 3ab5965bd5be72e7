"""Host-side logic of the training-loop layer: exponential families against golden vectors made by the
unmodified reference (tests/golden/make_golden_expfam.py), the EFN and APT losses against their closed forms.  No GPU."""
import os

import numpy as np
import pytest
import torch

from torch_nf_b200 import exponential_families as ef
from torch_nf_b200 import lfi
from torch_nf_b200.bijectors import ToSimplex

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "expfam.npz"))


@pytest.mark.parametrize("D", [2, 3, 5])
def test_mvn_matches_reference(D):
    fam = ef.MVN(D)
    assert fam.D_eta == int(G["mvn%d_D_eta" % D]) and fam.support_layer is None
    np.random.seed(10 + D)                      # same numpy stream consumption as the reference's sample_eta
    eta = fam.sample_eta(N=4)
    np.testing.assert_allclose(eta, G["mvn%d_eta" % D], rtol=1e-10, atol=1e-12)
    mu, Sigma = fam.eta_to_mu(G["mvn%d_eta" % D])
    np.testing.assert_allclose(mu, G["mvn%d_mu" % D], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(Sigma, G["mvn%d_Sigma" % D], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(fam.mu_to_eta(mu, Sigma), G["mvn%d_eta_back" % D], rtol=1e-8, atol=1e-10)
    T = fam.T(torch.tensor(G["mvn%d_z" % D]))
    assert T.shape == (2, 6, fam.D_eta)
    np.testing.assert_array_equal(T.numpy(), G["mvn%d_T" % D])
    np.testing.assert_allclose(fam.KL(G["mvn%d_kl_z" % D], G["mvn%d_kl_lp" % D], G["mvn%d_eta" % D]), G["mvn%d_KL" % D], rtol=1e-9)


def test_mvn_single_eta():
    np.random.seed(3)
    np.testing.assert_allclose(ef.MVN(3).sample_eta(N=1), G["mvn_eta_N1"], rtol=1e-10)


@pytest.mark.parametrize("D", [3, 4])
def test_dirichlet_matches_reference(D):
    fam = ef.Dirichlet(D)
    assert fam.D_eta == int(G["dir%d_D_eta" % D]) and fam.support_layer is ToSimplex
    np.random.seed(20 + D)
    np.testing.assert_array_equal(fam.sample_eta(N=5), G["dir%d_eta" % D])
    np.testing.assert_array_equal(fam.eta_to_mu(G["dir%d_eta" % D]), G["dir%d_alpha" % D])
    np.testing.assert_array_equal(fam.mu_to_eta(G["dir%d_alpha" % D]), G["dir%d_eta" % D])
    np.testing.assert_allclose(fam.T(torch.tensor(G["dir%d_z" % D])).numpy(), G["dir%d_T" % D], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(fam.KL(G["dir%d_kl_z" % D], G["dir%d_kl_lp" % D], G["dir%d_eta" % D]), G["dir%d_KL" % D], rtol=1e-9)


def test_family_validation():
    with pytest.raises(TypeError):
        ef.MVN(2.0)
    with pytest.raises(ValueError):
        ef.MVN(0)
    with pytest.raises(TypeError):
        ef.ExponentialFamily(3, support_layer=ToSimplex(3))      # the reference takes the CLASS, not an instance
    base = ef.ExponentialFamily(3)
    for call in (lambda: base.sample_eta(2), lambda: base.mu_to_eta(None), lambda: base.eta_to_mu(None), lambda: base.T(None)):
        with pytest.raises(NotImplementedError):
            call()


def test_T_is_differentiable():
    z = torch.randn(2, 3, 4, requires_grad=True)
    ef.MVN(4).T(z).sum().backward()
    # d/dz_k of sum_i z_i + sum_{i<=j} z_i z_j = 1 + z_k + sum_j z_j
    np.testing.assert_allclose(z.grad.numpy(), (1 + z + z.sum(dim=2, keepdim=True)).detach().numpy(), rtol=1e-5, atol=1e-6)


def test_efn_loss_closed_form():
    rs = np.random.RandomState(0)
    fam = ef.MVN(3)
    z = torch.tensor(rs.standard_normal((4, 5, 3)).astype(np.float32))
    lp = torch.tensor(rs.standard_normal((4, 5)).astype(np.float32))
    eta = torch.tensor(rs.standard_normal((4, fam.D_eta)).astype(np.float32))
    want = np.mean(lp.numpy() - np.einsum("mnk,mk->mn", fam.T(z).numpy(), eta.numpy()))
    assert abs(float(lfi.EFNLoss(z, lp, eta, fam.T)) - want) < 1e-5


class _FakeCNF(object):
    """log q(z | x) = -|z - x|^2 / 2 (a unit Gaussian around the context), differentiable in nothing."""

    def log_prob(self, z, x):
        return -0.5 * ((z - x[:, None, :]) ** 2).sum(dim=2)


def test_apt_atoms_and_loss():
    g = torch.Generator().manual_seed(0)
    M, A = 16, 5
    idx = lfi.apt_atoms(M, A, g)
    assert idx.shape == (M, A) and torch.equal(idx[:, 0], torch.arange(M))
    for m in range(M):
        row = idx[m].tolist()
        assert len(set(row)) == A and all(0 <= v < M for v in row)         # distinct, own parameter only in column 0
    with pytest.raises(ValueError):
        lfi.apt_atoms(4, 5)
    with pytest.raises(ValueError):
        lfi.apt_atoms(4, 1)
    rs = np.random.RandomState(1)
    z = torch.tensor(rs.standard_normal((M, 3)).astype(np.float32))
    x = torch.tensor(rs.standard_normal((M, 3)).astype(np.float32))
    loss = float(lfi.apt_loss(_FakeCNF(), z, x, A, atoms=idx))
    lp = -0.5 * ((z.numpy()[idx.numpy()] - x.numpy()[:, None, :]) ** 2).sum(axis=2)
    want = -np.mean(lp[:, 0] - np.log(np.exp(lp).sum(axis=1)))
    assert abs(loss - want) < 1e-5
    # a non-flat prior enters as q / p
    log_prior = lambda zz: -0.5 * (zz ** 2).sum(dim=-1)      # noqa: E731
    loss_p = float(lfi.apt_loss(_FakeCNF(), z, x, A, log_prior=log_prior, atoms=idx))
    lp2 = lp + 0.5 * (z.numpy()[idx.numpy()] ** 2).sum(axis=2)
    assert abs(loss_p - (-np.mean(lp2[:, 0] - np.log(np.exp(lp2).sum(axis=1))))) < 1e-5


def test_clip_grads():
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.tensor([-3.0, -0.5, 0.5, 3.0])
    q = torch.nn.Parameter(torch.zeros(2))              # no gradient: skipped
    lfi.clip_grads([p, q], 1.0)
    assert p.grad.tolist() == [-1.0, -0.5, 0.5, 1.0]
