"""Per-kernel device time of one C4 training step (torch.profiler, CUDA activities).  python profiles/scripts/train_breakdown.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch_nf_b200.density_estimator as de
from torch_nf_b200 import train
from torch_nf_b200.bijectors import ToInterval
from torch_nf_b200.conditional_density_estimator import ConditionalDensityEstimator
from torch.profiler import profile, ProfilerActivity
D, Dx, M = 6, 2, 1 << 18
dev = torch.device("cuda", 0)
np.random.seed(0); torch.manual_seed(0)
nf = de.NormFlow(D, True, sys.argv[1] if len(sys.argv) > 1 else "coupling", 1, 2, 15, ToInterval(D, [-2.0] * D, [2.0] * D))
cde = ConditionalDensityEstimator(nf, Dx, [64, 64]).to(dev)
opt = torch.optim.Adam(cde.parameters(), lr=1e-4)
x = torch.randn(M, Dx, device=dev); z = torch.rand(M, D, device=dev) * 3.8 - 1.9
step = lambda: train.train_step(lambda: train.nde_loss(cde, z, x), cde.parameters(), opt)
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
