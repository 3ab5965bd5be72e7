"""C4: conditional-flow training step (forward + backward + gradient all-reduce + Adam) for the scripts/lfi_mat.py
posterior shapes: NormFlow(D=6, conditioner, 'coupling', 1 stage, L=2, U=15, ToInterval) + hyper-network 2 -> 64 -> 64
-> D_params, M = 2^18 contexts per GPU, N = 1.  One JSON line (rank 0).
    python profiles/scripts/bench_train.py [--steps K] [--m 262144]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/scripts/bench_train.py"""
import argparse, json, os, sys
import numpy as np, torch
import torch.distributed as td
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch_nf_b200.density_estimator as de
from torch_nf_b200 import _lib, dist, train
from torch_nf_b200.bijectors import ToInterval
from torch_nf_b200.conditional_density_estimator import ConditionalDensityEstimator

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--m", type=int, default=1 << 18)
ap.add_argument("--arch", default="coupling", choices=["coupling", "AR"])
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    td.init_process_group("nccl", device_id=dev)
    dist.enable()
D, Dx, M = 6, 2, args.m
np.random.seed(0); torch.manual_seed(0)                       # identical replicas on every rank
nf = de.NormFlow(D, True, args.arch, 1, 2, max(15, 2 * D), ToInterval(D, [-2.0] * D, [2.0] * D))
cde = ConditionalDensityEstimator(nf, Dx, [64, 64]).to(dev)
opt = torch.optim.Adam(cde.parameters(), lr=1e-4)
g = torch.Generator(device=dev).manual_seed(100 + rank)       # this rank's shard of the contexts
x = torch.randn(M, Dx, device=dev, generator=g)
z = (torch.rand(M, D, device=dev, generator=g) * 3.8 - 1.9)   # inside the support (lb+0.1, ub-0.1)


def step():
    return train.train_step(lambda: train.nde_loss(cde, z, x), cde.parameters(), opt)


def barrier():
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()


losses = []
for _ in range(args.warmup):
    losses.append(step())
barrier()
l0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    losses.append(step())
e1.record(); barrier()
ms = e0.elapsed_time(e1)
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    ms = float(t.item())
losses = [float(v) for v in losses]
if rank == 0:
    P = nf.D_params
    # HBM bytes per context-step: the parameter row written by the hyper-network, read by forward, re-read by backward,
    # its gradient written by backward and read by the hyper-network's backward GEMMs (+ small z / activations)
    byts = 5 * P * 4
    print(json.dumps({
        "workload": "C4: conditional flow training step (fwd + bwd + grad all-reduce + Adam), arch %s, D=%d, D_params=%d, hidden [64,64], M=%d per GPU, N=1" % (args.arch, D, P, M),
        "value": world * M * args.steps / (ms * 1e-3), "unit": "samples/s", "n_gpus": world, "ms_per_step": ms / args.steps,
        "steps": args.steps, "warmup": args.warmup, "dtype": "fp32", "kernel_launches_per_step": (_lib.launch_count() - l0) // args.steps,
        "roofline": {"bound": "hbm (per-context parameter rows)", "bytes_per_sample": byts,
                     "hbm_gbs": byts * M * args.steps / (ms * 1e-3) / 1e9},
        "loss_first": losses[0], "loss_last": losses[-1], "finite": bool(np.isfinite(losses).all())}))
if world > 1:
    td.destroy_process_group()
