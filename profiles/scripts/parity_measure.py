"""Print the measured deviations from the reference goldens / the oracle that the stated tolerances are set against."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import flow_oracle as O
import torch_nf_b200 as tnf
import torch_nf_b200.density_estimator as de
from torch_nf_b200.synthetic import chain_spec, synthetic_params
T = torch.tensor
G = lambda n: np.load(os.path.join(ROOT, "tests", "golden", n + ".npz"))
rel = lambda a, b: float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max())
for name in ("flow_c3", "flow_c5"):
    g = G(name)
    D, stages, L, U, M, N, pseed, oseed = [int(v) for v in g["cfg"]]
    for mode in ("fp32_cc", "fp32", "bf16"):
        tnf.set_conditioner_precision(mode)
        nf = de.NormFlow(D, True, "coupling", stages, L, U)
        params = T(synthetic_params(chain_spec(nf.bijectors), D, M, seed=pseed)).cuda()
        np.random.seed(oseed)
        omega = np.random.normal(0.0, 1.0, (M, N, D))
        with torch.no_grad():
            z, lq = nf.forward(params, N, omega=omega)
            lp = nf.log_prob(T(g["z"]).cuda(), params)
        print("%s %s N=%d: max|dz|=%.3g rel_z=%.3g rel_logq=%.3g rel_logp=%.3g abs_logp=%.3g" % (
            name, mode, N, np.abs(z.cpu().numpy() - g["z"]).max(), rel(z.cpu().numpy(), g["z"]), rel(lq.cpu().numpy(), g["log_q_z"]),
            rel(lp.cpu().numpy(), g["log_prob"]), np.abs(lp.cpu().numpy() - g["log_prob"]).max()), flush=True)
# C3 / C5 against the oracle at larger N
for (D, stages, L, U, N) in ((64, 4, 2, 256, 1 << 16), (256, 8, 2, 256, 1 << 13)):
    chain = O.build_chain(D, "coupling", stages, L, U)
    nf = de.NormFlow(D, True, "coupling", stages, L, U)
    params = T(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=0))
    omega = np.random.RandomState(5).standard_normal((1, N, D))
    with torch.no_grad():
        zo, lqo, st = O.normflow_forward(chain, D, params, omega)
        lpo = O.normflow_log_prob(chain, D, zo, params, st)
    for mode in ("fp32_cc", "fp32", "bf16"):
        tnf.set_conditioner_precision(mode)
        with torch.no_grad():
            z, lq = nf.forward(params.cuda(), N, omega=omega)
            lp = nf.log_prob(zo.cuda(), params.cuda())
        print("oracle D=%d stages=%d N=%d %s: max|dz|=%.3g rel_z=%.3g rel_logq=%.3g rel_logp=%.3g abs_logp=%.3g" % (
            D, stages, N, mode, (z.cpu() - zo).abs().max(), rel(z.cpu().numpy(), zo.numpy()), rel(lq.cpu().numpy(), lqo.numpy()),
            rel(lp.cpu().numpy(), lpo.numpy()), (lp.cpu() - lpo).abs().max()), flush=True)
tnf.set_conditioner_precision("fp32")
