"""Tensor-core coupling backward (tnf_coupling_tc_bwd): kernel time, weight-gradient GEMM time, and the C3 maximum-
likelihood training step (-mean log_prob, forward + backward + Adam) in the bf16-conditioner mode.
Usage: python profiles/scripts/tcb_bench.py [--rows 1048576] [--cc-rows 16384]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch_nf_b200.density_estimator as de   # noqa: E402
from torch_nf_b200 import config, ops, _lib   # noqa: E402
from torch_nf_b200._lib import TNF_INVERSE   # noqa: E402
from torch_nf_b200.synthetic import chain_spec, synthetic_params   # noqa: E402


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 20)
    ap.add_argument("--cc-rows", type=int, default=1 << 14)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda")
    D, U, L, rows = 64, 256, 2, a.rows
    out = {}
    params = torch.tensor(synthetic_params([("RealNVP", L, U, True)], D, 1, seed=0)).to(dev)
    z = torch.randn(1, rows, D, device=dev)
    gz = torch.randn(1, rows, D, device=dev)
    gl = torch.randn(rows, device=dev)
    packed = ops.tc_bwd_pack(params[0], D, U, L, True)
    # ---- the kernel alone (ctypes call, preallocated workspace)
    lib = _lib.lib()
    ws = torch.empty(lib.tnf_tc_bwd_workspace_bytes(rows, D, U, L) // 2, dtype=torch.bfloat16, device=dev)
    g_z = torch.empty_like(z)
    st = torch.cuda.current_stream().cuda_stream

    def kern():
        rc = lib.tnf_coupling_tc_bwd(z.data_ptr(), packed.data_ptr(), gz.data_ptr(), gl.data_ptr(), g_z.data_ptr(),
                                     ws.data_ptr(), rows, D, U, L, 1, TNF_INVERSE, 0, 0, st)
        assert rc == 0
    ms_k = timed(kern)
    flop = 3 * 327680 * rows        # recompute (1x) + data gradient (1x) on the kernel; weight gradient (1x) in the GEMMs
    out["kernel_ms"] = ms_k
    out["kernel_tflops_recompute_plus_dgrad"] = 2 * 327680 * rows / ms_k / 1e9
    out["workspace_GB"] = ws.numel() * 2 / 1e9
    out["kernel_hbm_GBps"] = (ws.numel() * 2 + 3 * rows * D * 4) / ms_k / 1e6
    g_params = torch.zeros_like(params)
    ms_all = timed(lambda: ops.coupling_tc_bwd(z, packed, gz, gl, g_params[0], D, U, L, True, TNF_INVERSE), n=5)
    out["layer_backward_ms"] = ms_all
    out["weight_gradient_gemms_ms"] = ms_all - ms_k
    out["layer_backward_tflops_algorithmic"] = flop / ms_all / 1e9
    # ---- CUDA-core backward of the same layer (the path it replaces), on fewer rows
    rc_ = a.cc_rows
    gp2 = torch.zeros_like(params)
    ms_cc = timed(lambda: ops.coupling_bwd(z[:, :rc_], params, gz[:, :rc_], gl[:rc_].view(1, -1), gp2, D, U, L, True, TNF_INVERSE), n=3, warm=1)
    out["cuda_core_backward_ms_per_2^20_rows"] = ms_cc * rows / rc_
    del ws, g_z
    # ---- C3 maximum-likelihood training step
    config.set_conditioner_precision("bf16")
    nf = de.NormFlow(D, False, "coupling", 4, 2, U)
    p = torch.tensor(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=0)).to(dev).requires_grad_(True)
    nf.params = p
    with torch.no_grad():
        zs, _ = nf.forward(p.detach(), rows)
    zs = zs.detach()
    opt = torch.optim.Adam([p], lr=1e-4)
    losses = []

    def step():
        opt.zero_grad(set_to_none=True)
        loss = -nf.log_prob(zs).mean()
        loss.backward()
        opt.step()
        losses.append(loss.detach())
    l0 = _lib.launch_count()
    ms_step = timed(step, n=a.steps, warm=2)
    out["c3_train_step_ms"] = ms_step
    out["c3_train_samples_per_s"] = rows / ms_step * 1e3
    out["c3_train_launches_per_step"] = (_lib.launch_count() - l0) / (a.steps + 2)
    out["c3_train_losses"] = [float(x) for x in losses]
    out["c3_train_tflops_algorithmic"] = 8 * 4 * 327680 * rows / ms_step / 1e9     # forward + 3x backward per layer
    out["rows"] = rows
    print(json.dumps(out))


if __name__ == "__main__":
    main()
