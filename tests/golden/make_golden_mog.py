"""Golden vectors of the mixture-of-Gaussians estimator from the UNMODIFIED reference (torch_nf/density_estimator.py:57-237).
Run in the build container only (needs /root/reference):  python tests/golden/make_golden_mog.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import torch_nf.density_estimator as rde   # noqa: E402  (the reference)

out = {}
rs = np.random.RandomState(7)
for tag, D, K, bounded in (("k1", 3, 1, False), ("k3", 4, 3, False), ("k2b", 3, 2, True), ("k1b", 2, 1, True)):
    lb = -2.0 - 0.1 * np.arange(D) if bounded else None
    ub = 3.0 + 0.2 * np.arange(D) if bounded else None
    mog = rde.MoG(D, True, K, lb=lb, ub=ub)
    M, N = 5, 7
    params = torch.tensor(rs.standard_normal((M, mog.D_params)).astype(np.float32) * 0.7)
    alpha, mu, Sigma_inv, Sigma_det = mog._get_MoG_params(params)
    z = torch.tensor(rs.standard_normal((M, N, D)).astype(np.float32) * 1.5)
    out[tag + "_D"], out[tag + "_K"], out[tag + "_D_params"] = D, K, mog.D_params
    if bounded:
        out[tag + "_lb"], out[tag + "_ub"] = lb, ub
    out[tag + "_params"] = params.numpy()
    out[tag + "_alpha"], out[tag + "_mu"] = alpha.numpy(), mu.numpy()
    out[tag + "_Sigma_inv"], out[tag + "_Sigma_det"] = Sigma_inv.numpy(), Sigma_det.numpy()
    out[tag + "_z"] = z.numpy()
    out[tag + "_log_prob"] = mog.log_prob(z, params).numpy()
    out[tag + "_log_prob_np"] = mog.log_prob_np(z.numpy().astype(np.float64), params)
np.savez_compressed(os.path.join(HERE, "mog.npz"), **out)
print("wrote mog.npz", len(out), "arrays")
