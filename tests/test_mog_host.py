"""MoG: constructor / validation / parameter map against golden vectors made by the unmodified reference
(tests/golden/make_golden_mog.py).  The parameter map runs on the device of ``params`` (here the CPU); sampling and
densities need the GPU (tests/test_gpu_mog.py)."""
import os

import numpy as np
import pytest
import torch

import torch_nf_b200.density_estimator as de
from torch_nf_b200.conditional_density_estimator import ConditionalDensityEstimator

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "mog.npz"))
TAGS = ("k1", "k3", "k2b", "k1b")


def make(tag):
    D, K = int(G[tag + "_D"]), int(G[tag + "_K"])
    lb = G[tag + "_lb"] if tag + "_lb" in G else None
    ub = G[tag + "_ub"] if tag + "_ub" in G else None
    return de.MoG(D, True, K, lb=lb, ub=ub)


def test_mog_ctor_and_validation():
    mog = de.MoG(4, False, 1)
    assert mog.D == 4 and mog.K == 1 and not mog.conditioner and mog.D_params == 1 + 4 + 10
    assert tuple(mog.params.shape) == (1, 15) and mog.params.requires_grad
    with pytest.raises(TypeError):
        de.MoG(4, True, 2.)
    with pytest.raises(ValueError):
        de.MoG(4, True, 0)
    cde = ConditionalDensityEstimator(de.MoG(4, True, 3), 10, [50, 50])       # accepted next to NormFlow (:48)
    assert cde.D_params == 3 * (1 + 4 + 10)


@pytest.mark.parametrize("tag", TAGS)
def test_mog_parameter_map_matches_reference(tag):
    mog = make(tag)
    assert mog.D_params == int(G[tag + "_D_params"])
    alpha, mu, Sigma_inv, Sigma_det = mog._get_MoG_params(torch.tensor(G[tag + "_params"]))
    np.testing.assert_allclose(alpha.numpy(), G[tag + "_alpha"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(mu.numpy(), G[tag + "_mu"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(Sigma_inv.numpy(), G[tag + "_Sigma_inv"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(Sigma_det.numpy(), G[tag + "_Sigma_det"], rtol=1e-5, atol=1e-12)
    a_np, mu_np, P_np, _ = mog._get_MoG_params(torch.tensor(G[tag + "_params"]), numpy=True)
    assert isinstance(a_np, np.ndarray) and np.allclose(a_np.sum(axis=1), 1.0) and P_np.shape == Sigma_inv.shape
