"""Time one C3 coupling layer (D=64, U=256, L=2, 2^20 rows) in every mode the chain uses it in: inverse, forward,
forward + fused BatchNorm statistics (incl. the statistics reduce launch), with a pre-affine, and the whole chain's
two directions.   python profiles/scripts/tc_modes_bench.py [precision]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch_nf_b200 as tnf
import torch_nf_b200.density_estimator as de
from torch_nf_b200 import ops
from torch_nf_b200.synthetic import synthetic_params, chain_spec

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
D, U, L, N = 64, 256, 2, 1 << 20
params = torch.tensor(synthetic_params([("RealNVP", L, U, True)], D, 1, seed=0))
packed = ops.tc_pack(params.cuda()[0], D, U, L, True, precision=prec)
torch.manual_seed(0)
z = torch.randn(1, N, D, device="cuda")
out = torch.empty_like(z.reshape(-1, D))
ps = torch.rand(D, device="cuda") + 0.5
pb = torch.randn(D, device="cuda")


def timeit(name, fn, reps=3, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n)
    print("%-44s %.4f ms (min of %s)" % (name, min(ts), ",".join("%.4f" % t for t in ts)), flush=True)


timeit("inverse", lambda: ops.coupling_tc(z, packed, D, U, L, True, ops.TNF_INVERSE, out=out, precision=prec))
timeit("forward", lambda: ops.coupling_tc(z, packed, D, U, L, True, ops.TNF_FORWARD, out=out, precision=prec))
timeit("forward + pre-affine", lambda: ops.coupling_tc(z, packed, D, U, L, True, ops.TNF_FORWARD, out=out, pre_scale=ps, pre_shift=pb, precision=prec))
if prec == "bf16":
    timeit("forward + statistics (+ reduce launch)", lambda: ops.coupling_tc(z, packed, D, U, L, True, ops.TNF_FORWARD, out=out, want_stats=True, precision=prec))
tnf.set_conditioner_precision(prec)
nf = de.NormFlow(D, True, "coupling", 4, L, U)
cp = torch.tensor(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=0)).cuda()
with torch.no_grad():
    zz, lq = nf.forward(cp, N)
    timeit("chain sample (8 layers)", lambda: nf.forward(cp, N), n=5)
    timeit("chain log_prob (8 layers)", lambda: nf.log_prob(zz, cp), n=5)
    timeit("base_sample", lambda: ops.base_sample(1, N, D, 1, 0, zz.device), n=10)
    sc = (torch.ones(D, device="cuda"), torch.zeros(D, device="cuda"))
    timeit("colaffine", lambda: ops.colaffine(zz, sc, D), n=10)
    timeit("colstats", lambda: ops.colstats(zz, D), n=10)
