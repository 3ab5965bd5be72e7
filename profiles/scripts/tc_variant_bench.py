"""Time one tensor-core coupling layer (C3 shape by default) for a list of `variant` values of tnf_coupling_tc and
compare each against variant 0 and against the bf16-emulating oracle.
    python profiles/scripts/tc_variant_bench.py 0 256 512 ...      (variant = kernel | tune << 8)"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import flow_oracle as O
from torch_nf_b200 import ops
from torch_nf_b200.synthetic import synthetic_params

D, U, L, N = 64, 256, 2, 1 << 20
variants = [int(v) for v in sys.argv[1:]] or [0]
params = torch.tensor(synthetic_params([("RealNVP", L, U, True)], D, 1, seed=0))
packed = ops.tc_pack(params.cuda()[0], D, U, L, True)
torch.manual_seed(0)
z = torch.randn(1, N, D, device="cuda")
ns = 2048
ze, lde = O.coupling_bf16_emulated(z[:, :ns].cpu(), params, D, L, U, True, True)
zr, ldr = O.coupling_inverse(z[:, :ns].cpu(), params, D, L, U, True)
base = None
for v in variants:
    for _ in range(3):
        zo, ld = ops.coupling_tc(z, packed, D, U, L, True, ops.TNF_INVERSE, variant=v)
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            zo, ld = ops.coupling_tc(z, packed, D, U, L, True, ops.TNF_INVERSE, variant=v)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 10)
    zf, ldf = ops.coupling_tc(z, packed, D, U, L, True, ops.TNF_FORWARD, variant=v & 0xFFF)
    torch.cuda.synchronize()
    if base is None:
        base = (zo.clone(), ld.clone(), zf.clone())
    print("variant %6d: %.4f ms/launch (min of 3x10: %s)  vs v0: max|dz|=%.3g max|dld|=%.3g fwd %.3g | vs emulated: %.3g / %.3g | vs fp32: %.3g / %.3g | finite %s" % (
        v, min(ts), ",".join("%.4f" % t for t in ts), (zo - base[0]).abs().max().item(), (ld - base[1]).abs().max().item(),
        (zf - base[2]).abs().max().item(),
        (zo[:, :ns].cpu() - ze).abs().max().item(), (ld[:ns].cpu().view(1, ns) - lde).abs().max().item(),
        (zo[:, :ns].cpu() - zr).abs().max().item(), (ld[:ns].cpu().view(1, ns) - ldr).abs().max().item(),
        bool(torch.isfinite(zo).all())), flush=True)
