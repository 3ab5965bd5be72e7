"""Device time of the fused conditional log-density kernel alone (tnf_cde_logprob, both producers), without the Python
call path and the torch hidden layers: C4-like (D=6, H=64, M=2^18) and C2b (D=8, H=100, M=2^16).
    python profiles/scripts/cde_kernel_bench.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch_nf_b200.density_estimator as de
from torch_nf_b200 import _lib, ops
from torch_nf_b200.bijectors import ToInterval
from torch_nf_b200.conditional_density_estimator import ConditionalDensityEstimator, _NoParams

lib = _lib.lib()
for name, D, Dx, hidden, M, sup in (("c4like", 6, 2, [64, 64], 1 << 18, True), ("c2b", 8, 8, [100], 1 << 16, False)):
    torch.manual_seed(0)
    nf = de.NormFlow(D, True, "coupling", 1, 2, 15, ToInterval(D, [-2.0] * D, [2.0] * D) if sup else None)
    cde = ConditionalDensityEstimator(nf, Dx, hidden).cuda()
    x = torch.randn(M, Dx, device="cuda")
    z = (torch.rand(M, 1, D, device="cuda") * 3.6 - 1.8) if sup else torch.randn(M, 1, D, device="cuda")
    dev = x.device
    arr, keep, _ = nf._chain_pod(_NoParams(dev, M), de._Rows(M, 0, torch.float32), None, sample=False)
    last = cde.param_net[-1]
    H = last.in_features
    with torch.no_grad():
        h = cde.param_net[:-1](x).contiguous()
    lp = torch.empty(M, device=dev)
    extra = [(0 | (int(b) << 8), "tcgen05 dbg%s" % b) for b in os.environ.get("TNF_CDE_DBG", "").split(",") if b]
    for variant, vname in [(0, "tcgen05"), (1, "cuda cores")] + extra:
        packed = torch.empty(lib.tnf_cde_packed_bytes(nf.D_params, H, variant & 15), dtype=torch.uint8, device=dev)
        _lib.check(lib.tnf_cde_pack(arr, len(nf.bijectors), D, last.weight.data_ptr(), last.bias.data_ptr(), H, packed.data_ptr(), variant & 15, ops._stream()), "pack")
        call = lambda: _lib.check(lib.tnf_cde_logprob(arr, len(nf.bijectors), D, h.data_ptr(), H, packed.data_ptr(), z.data_ptr(), M, lp.data_ptr(), variant, ops._stream()), "lp")
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                call()
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 20)
        t = min(ts)
        if (variant >> 8) & 4:      # timing experiment: clock64 stamps of CTA 0 (MMA warp: 4 per block; consumer thread 0: 2 per block)
            lp.zero_(); call(); torch.cuda.synchronize()
            raw = lp.view(torch.int64).cpu().numpy()
            mm = raw[:4 * 40].reshape(-1, 4); cs = raw[512:512 + 2 * 40].reshape(-1, 2)
            t0 = mm[0, 0]
            print("block: MMA warp [start, empty seen, W seen, issued+committed] | consumer [wait full from, seen]   (cycles since the first stamp)")
            for i in range(26):
                print("%3d: %7d %7d %7d %7d | %7d %7d" % ((i,) + tuple(int(v - t0) for v in mm[i]) + tuple(int(v - t0) for v in cs[i])))
        print("%-7s %-10s %.4f ms/launch  %.3g samples/s  algorithmic %.1f TFLOP/s (2 (H+1) D_params per sample)  h+z+lp traffic %.0f GB/s" % (
            name, vname, t, M / (t * 1e-3), 2.0 * (H + 1) * nf.D_params * M / (t * 1e-3) / 1e12, 4.0 * (H + D + 1) * M / (t * 1e-3) / 1e9), flush=True)
