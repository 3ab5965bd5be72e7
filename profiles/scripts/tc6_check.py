"""fp32-parity tensor-core coupling layer (precision="fp32_tc") against the fp32 oracle; timing at the C3 shape."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import flow_oracle as O
from torch_nf_b200 import ops
from torch_nf_b200.synthetic import synthetic_params
T = torch.tensor
for (D, U, L, N, upper) in ((64, 256, 2, 1000, True), (64, 256, 2, 333, False), (64, 128, 2, 257, True), (128, 256, 2, 384, True),
                            (64, 256, 1, 300, True), (64, 256, 3, 500, False), (128, 128, 2, 129, False), (256, 256, 2, 300, True), (256, 128, 2, 200, False)):
    params = T(synthetic_params([("RealNVP", L, U, upper)], D, 1, seed=3))
    z = torch.randn(1, N, D, generator=torch.Generator().manual_seed(1)) * 1.3
    packed = ops.tc_pack(params.cuda()[0], D, U, L, upper, precision="fp32_tc")
    for direction, fn in ((ops.TNF_FORWARD, O.coupling_forward), (ops.TNF_INVERSE, O.coupling_inverse)):
        zo, ldo = fn(z, params, D, L, U, upper)
        zd, ld = ops.coupling_tc(z.cuda(), packed, D, U, L, upper, direction, precision="fp32_tc")
        torch.cuda.synchronize()
        # float64 reference of the same layer for the oracle's own fp32 noise
        z64, ld64 = fn(z.double(), params.double(), D, L, U, upper)
        rz = ((zd.cpu() - zo).abs() / zo.abs().clamp(min=1)).max().item()
        rl = ((ld.cpu().view(1, N) - ldo).abs() / ldo.abs().clamp(min=1)).max().item()
        rz64 = ((zd.cpu().double() - z64).abs() / z64.abs().clamp(min=1)).max().item()
        oz64 = ((zo.double() - z64).abs() / z64.abs().clamp(min=1)).max().item()
        print("D=%d U=%d L=%d N=%d upper=%d dir=%d: rel_z=%.3g rel_ld=%.3g | vs f64: ours %.3g, torch fp32 %.3g" % (
            D, U, L, N, upper, direction, rz, rl, rz64, oz64), flush=True)
D, U, L, N = 64, 256, 2, 1 << 20
params = T(synthetic_params([("RealNVP", L, U, True)], D, 1, seed=0))
packed = ops.tc_pack(params.cuda()[0], D, U, L, True, precision="fp32_tc")
z = torch.randn(1, N, D, device="cuda")
for _ in range(2):
    ops.coupling_tc(z, packed, D, U, L, True, ops.TNF_INVERSE, precision="fp32_tc")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.coupling_tc(z, packed, D, U, L, True, ops.TNF_INVERSE, precision="fp32_tc")
e1.record(); torch.cuda.synchronize()
print("C3 layer, 2^20 rows, fp32_tc: %.4f ms per launch" % (e0.elapsed_time(e1) / 5))
for prec in ("bf16", "fp32_tc"):
    D, U, L, N = 256, 256, 2, 1 << 19
    params = T(synthetic_params([("RealNVP", L, U, True)], D, 1, seed=0))
    packed = ops.tc_pack(params.cuda()[0], D, U, L, True, precision=prec)
    z = torch.randn(1, N, D, device="cuda")
    for _ in range(2):
        ops.coupling_tc(z, packed, D, U, L, True, ops.TNF_INVERSE, precision=prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        zo, ld = ops.coupling_tc(z, packed, D, U, L, True, ops.TNF_INVERSE, precision=prec)
    e1.record(); torch.cuda.synchronize()
    ze, lde = O.coupling_bf16_emulated(z[:, :512].cpu(), params, D, L, U, True, True)
    print("C5 layer (D=256), 2^19 rows, %s: %.4f ms per launch (%.0f TFLOP/s algorithmic); vs bf16-emulated oracle max|dz| %.3g" % (
        prec, e0.elapsed_time(e1) / 5, 524288.0 * N / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e12, (zo[:, :512].cpu() - ze).abs().max().item()))
