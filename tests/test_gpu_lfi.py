"""The training loops of torch_nf_b200/lfi.py on the GPU: a few iterations of SNPE, APT and EFN training on toy systems
(losses finite, SNPE / EFN losses decrease, return shapes of scripts/lfi_mat.py:48-57)."""
import numpy as np
import pytest
import torch

import torch_nf_b200.density_estimator as de
from torch_nf_b200 import exponential_families as ef
from torch_nf_b200 import lfi
from torch_nf_b200.bijectors import ToInterval
from torch_nf_b200.conditional_density_estimator import ConditionalDensityEstimator

pytestmark = pytest.mark.gpu


class GaussToy(object):
    """x = z + 0.1 noise, uniform prior on [-2, 2]^D (the System protocol of LFI_learning_rules.ipynb:203-213)."""

    def __init__(self, D):
        self.D = D
        self.lb, self.ub = -2.0 * np.ones(D), 2.0 * np.ones(D)
        self.support_layer = ToInterval(D, self.lb, self.ub)
        self.rs = np.random.RandomState(0)

    def sample_prior(self, M):
        return self.rs.uniform(-1.9, 1.9, (M, self.D)), np.full((M,), 1.0 / 3.8 ** self.D)

    def simulate(self, z):
        return z + 0.1 * self.rs.standard_normal(z.shape)


def _cnf(D, D_x, support=None, seed=0):
    np.random.seed(seed); torch.manual_seed(seed)
    nf = de.NormFlow(D, True, "coupling", 1, 2, 15, support)
    return ConditionalDensityEstimator(nf, D_x, [32])


def test_train_snpe_and_apt():
    D = 3
    system = GaussToy(D)
    x0 = np.zeros((1, D))
    cnf = _cnf(D, D, system.support_layer)
    cnf, losses, zs, lps, it_time = lfi.train_SNPE(cnf, system, x0, M=256, R=2, num_iters=15, lr=1e-2)
    assert losses.shape == (30,) and np.isfinite(losses).all() and losses[-1] < losses[0]
    assert zs.shape == (2, 256, D) and lps.shape == (2, 256) and it_time > 0
    assert np.abs(zs).max() <= 2.0                       # samples stay inside the support layer's interval
    cnf = _cnf(D, D, system.support_layer, seed=1)
    cnf, losses, zs, lps, _ = lfi.train_APT(cnf, system, x0, M=128, M_atom=8, R=2, num_iters=10, lr=1e-2)
    assert losses.shape == (20,) and np.isfinite(losses).all()
    assert 0.0 <= losses.min() <= np.log(8) + 1e-3       # -log of a softmax weight among 8 atoms
    assert losses[-5:].mean() < losses[:5].mean()
    assert zs.shape == (2, 128, D) and lps.shape == (2, 128)


def test_train_efn():
    D = 2
    fam = ef.MVN(D)
    np.random.seed(0)
    cnf = _cnf(D, fam.D_eta)
    losses, KLs = lfi.train_efn(cnf, fam, num_iters=12, M=16, N=64, lr=1e-2)
    assert len(losses) == 12 and np.isfinite(losses).all() and np.isfinite(KLs).all()
    assert np.mean(losses[-4:]) < np.mean(losses[:4])
