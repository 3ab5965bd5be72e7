"""torch.profiler breakdown of the C3 maximum-likelihood training step (device time per kernel, host time per step)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch_nf_b200.density_estimator as de
from torch_nf_b200 import config
from torch_nf_b200.synthetic import chain_spec, synthetic_params
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda")
config.set_conditioner_precision("bf16")
D, U, rows = 64, 256, 1 << 20
nf = de.NormFlow(D, False, "coupling", 4, 2, U)
p = torch.tensor(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=0)).to(dev).requires_grad_(True)
nf.params = p
np.random.seed(0)
with torch.no_grad():
    nf.forward(p.detach(), 1 << 16)
z = (torch.randn(1, rows, D, device=dev) * 1.2 + 0.1).contiguous()
opt = torch.optim.Adam([p], lr=1e-4)


def step():
    opt.zero_grad(set_to_none=True)
    loss = -nf.log_prob(z).mean()
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
t_host = (time.perf_counter() - t0) / 5          # host time to ENQUEUE a step (no sync inside)
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 5
print("host enqueue per step %.2f ms, wall per step incl. device %.2f ms" % (t_host * 1e3, t_all * 1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
rows_ = []
for e in prof.key_averages():
    dt = getattr(e, "device_time_total", None)
    if dt is None:
        dt = getattr(e, "cuda_time_total", 0)
    if dt > 0 and e.device_type.name == "CUDA":
        rows_.append((dt / 3e3, e.count // 3, e.key[:90]))
rows_.sort(reverse=True)
tot = sum(r[0] for r in rows_)
print("device time per step %.2f ms" % tot)
for r in rows_[:25]:
    print("%8.3f ms  x%-3d %s" % r)
