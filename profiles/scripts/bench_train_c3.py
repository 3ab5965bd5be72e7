"""C3 maximum-likelihood training step on tensor cores: NormFlow(64, False, 'coupling', 4, 2, 256), loss = -mean log_prob(z)
over 2^20 samples per GPU (rows sharded over the ranks), forward (tnf_coupling_tc) + backward (tnf_coupling_tc_bwd +
weight-gradient GEMMs) of all 8 coupling layers, BatchNorm / Affine backward, ONE gradient all-reduce of the (1, D_params)
row, Adam.  bf16-conditioner mode.  One JSON line (rank 0).
    python profiles/scripts/bench_train_c3.py [--steps K] [--rows 1048576]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/scripts/bench_train_c3.py"""
import argparse, json, os, sys
import numpy as np, torch
import torch.distributed as td
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch_nf_b200.density_estimator as de
from torch_nf_b200 import _lib, config, dist, train
from torch_nf_b200.synthetic import chain_spec, synthetic_params

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--rows", type=int, default=1 << 20)
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
config.set_conditioner_precision("bf16")
D, U, STAGES, rows = 64, 256, 4, args.rows
nf = de.NormFlow(D, False, "coupling", STAGES, 2, U)
p = torch.tensor(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=0)).to(dev).requires_grad_(True)   # identical replicas
nf.params = p
np.random.seed(0)
with torch.no_grad():                       # the BatchNorm statistics log_prob uses: one forward, same noise on every rank
    nf.forward(p.detach(), 1 << 16)
if world > 1:
    td.init_process_group("nccl", device_id=dev)
    dist.enable()
g = torch.Generator(device=dev).manual_seed(100 + rank)       # this rank's shard of the "data" (not distributed as the model)
z = (torch.randn(1, rows, D, device=dev, generator=g) * 1.2 + 0.1).contiguous()
opt = torch.optim.Adam([p], lr=1e-4)


def step():
    return train.train_step(lambda: train.mle_loss(nf, z), [p], opt)


def barrier():
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()


losses = [step() for _ in range(args.warmup)]
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
    flop = 8 * 4 * 327680            # per sample: 8 coupling layers x (forward + recompute + data gradient + weight gradient)
    print(json.dumps({
        "workload": "C3 maximum-likelihood training step (fwd + bwd on tensor cores + grad all-reduce + Adam), NormFlow(64,False,'coupling',4,2,256), %d rows per GPU" % rows,
        "value": world * rows * args.steps / (ms * 1e-3), "unit": "samples/s", "n_gpus": world, "ms_per_step": ms / args.steps,
        "steps": args.steps, "warmup": args.warmup, "dtype": "bf16 conditioner (tensor cores), fp32 coupling / log-det / gradients accumulated in fp32",
        "kernel_launches_per_step": (_lib.launch_count() - l0) // args.steps,
        "roofline": {"bound": "tensor", "algorithmic_flop_per_sample": flop, "achieved": flop * rows * args.steps / (ms * 1e-3) / 1e12, "unit": "TFLOP/s per GPU"},
        "loss_first": losses[0], "loss_last": losses[-1], "finite": bool(np.isfinite(losses).all())}))
if world > 1:
    td.destroy_process_group()
