import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
import torch_nf_b200.density_estimator as de
from torch_nf_b200 import config
from torch_nf_b200.synthetic import chain_spec, synthetic_params
D, N = 8, 65536
nf = de.NormFlow(D, True, "coupling", 1, 2, 15)
params = torch.tensor(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=0)).cuda()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    z, lq = nf.forward(params, N)
    for chain in (True, False):
        config.set_chain_abi(chain)
        print("chain_abi", chain, "forward %.4f ms, log_prob %.4f ms, forward(freeze_bn) %.4f ms" % (t(lambda: nf.forward(params, N)), t(lambda: nf.log_prob(z, params)), t(lambda: nf.forward(params, N, freeze_bn=True))))
    config.set_chain_abi(True)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): nf.forward(params, N); nf.log_prob(z, params)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=6, max_name_column_width=60))
