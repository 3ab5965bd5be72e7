"""Golden vectors of the exponential families from the UNMODIFIED reference (torch_nf/exponential_families.py).
Run in the build container only (needs /root/reference):  python tests/golden/make_golden_expfam.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import torch_nf.exponential_families as ref   # noqa: E402  (the reference)

out = {}
rs = np.random.RandomState(5)
# MVN
for D in (2, 3, 5):
    fam = ref.MVN(D)
    np.random.seed(10 + D)
    eta = fam.sample_eta(N=4)
    mu, Sigma = fam.eta_to_mu(eta)
    z = rs.standard_normal((2, 6, D)).astype(np.float32)
    out["mvn%d_eta" % D] = eta
    out["mvn%d_mu" % D] = mu
    out["mvn%d_Sigma" % D] = Sigma
    out["mvn%d_eta_back" % D] = fam.mu_to_eta(mu, Sigma)
    out["mvn%d_z" % D] = z
    out["mvn%d_T" % D] = fam.T(torch.tensor(z)).numpy()
    out["mvn%d_D_eta" % D] = fam.D_eta
    lp = rs.standard_normal((4, 6))
    zz = rs.standard_normal((4, 6, D))
    out["mvn%d_kl_z" % D] = zz
    out["mvn%d_kl_lp" % D] = lp
    out["mvn%d_KL" % D] = fam.KL(zz, lp, eta)
np.random.seed(3)
out["mvn_eta_N1"] = ref.MVN(3).sample_eta(N=1)
# Dirichlet
for D in (3, 4):
    fam = ref.Dirichlet(D)
    np.random.seed(20 + D)
    eta = fam.sample_eta(N=5)
    z = rs.dirichlet(np.ones(D), size=(2, 7)).astype(np.float32)
    out["dir%d_eta" % D] = eta
    out["dir%d_alpha" % D] = fam.eta_to_mu(eta)
    out["dir%d_z" % D] = z
    out["dir%d_T" % D] = fam.T(torch.tensor(z)).numpy()
    out["dir%d_D_eta" % D] = fam.D_eta
    lp = rs.standard_normal((5, 7))
    zz = rs.dirichlet(np.ones(D), size=(5, 7))
    out["dir%d_kl_z" % D] = zz
    out["dir%d_kl_lp" % D] = lp
    out["dir%d_KL" % D] = fam.KL(zz, lp, eta)
np.savez_compressed(os.path.join(HERE, "expfam.npz"), **out)
print("wrote expfam.npz", len(out), "arrays")
