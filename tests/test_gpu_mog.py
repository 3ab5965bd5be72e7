"""MoG densities and sampling on the GPU against the reference's golden vectors and its own float64 mixture density."""
import os

import numpy as np
import pytest
import torch

from test_mog_host import G, TAGS, make

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", TAGS)
def test_mog_log_prob_matches_reference(tag):
    mog = make(tag)
    params, z = torch.tensor(G[tag + "_params"]), torch.tensor(G[tag + "_z"])
    lp = mog.log_prob(z, params)
    assert not lp.is_cuda and lp.dtype == torch.float32          # returned on the caller's device
    np.testing.assert_allclose(lp.numpy(), G[tag + "_log_prob"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(mog.log_prob_np(G[tag + "_z"].astype(np.float64), params), G[tag + "_log_prob_np"], rtol=1e-5, atol=1e-6)
    lp_dev = mog.log_prob(z.cuda(), params.cuda())
    assert lp_dev.is_cuda
    np.testing.assert_allclose(lp_dev.cpu().numpy(), lp.numpy(), rtol=1e-6, atol=1e-6)


def test_mog_sampling():
    """Samples follow the mixture (first two moments of a well separated K = 2 mixture) and log_q_z is the reference's
    float64 mixture density of the drawn points."""
    D, K, M, N = 2, 2, 3, 20000
    mog = __import__("torch_nf_b200.density_estimator", fromlist=["MoG"]).MoG(D, True, K)
    rs = np.random.RandomState(0)
    params = torch.tensor(rs.standard_normal((M, mog.D_params)).astype(np.float32) * 0.5)
    np.random.seed(3)
    z, lq = mog.forward(params, N=N)
    assert z.shape == (M, N, D) and lq.shape == (M, N) and z.dtype == torch.float32 and lq.dtype == torch.float32
    np.testing.assert_allclose(lq.numpy(), mog.log_prob_np(z.numpy().astype(np.float64), params), rtol=1e-4, atol=1e-4)
    alpha, mu, P, _ = mog._get_MoG_params(params, numpy=True)
    Sigma = np.linalg.inv(P) + 0.001 * np.eye(D)
    for m in range(M):
        mean = (alpha[m][:, None] * mu[m]).sum(axis=0)
        second = sum(alpha[m, k] * (Sigma[m, k] + np.outer(mu[m, k], mu[m, k])) for k in range(K))
        zm = z[m].numpy().astype(np.float64)
        assert np.abs(zm.mean(axis=0) - mean).max() < 0.05 * (1 + np.abs(mean).max())
        assert np.abs(zm.T @ zm / N - second).max() < 0.08 * (1 + np.abs(second).max())
    np.random.seed(3)
    z2, _ = mog.forward(params, N=N)
    assert torch.equal(z, z2)                                       # np.random.seed reproduces the draw, as for NormFlow
