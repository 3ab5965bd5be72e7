"""Device-resident throughput of the other BASELINE.json configurations (C3 is bench.py): one JSON line each.
    python profiles/scripts/bench_configs.py [c1 c2a c2b c5 c4like]
sample(N) + log_prob of the samples per step, CUDA events, 3 warm-ups + 10 steps, inputs resident in HBM."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch_nf_b200 as tnf
import torch_nf_b200.density_estimator as de
from torch_nf_b200 import _lib
from torch_nf_b200.bijectors import ToInterval
from torch_nf_b200.conditional_density_estimator import ConditionalDensityEstimator
from torch_nf_b200.synthetic import chain_spec, synthetic_params

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
T = torch.tensor


def timed(fn, steps=10, warmup=3):
    with torch.no_grad():
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (_lib.launch_count() - l0) // steps, out


def run(name):
    if name in ("c1", "c2a", "c3small", "c5"):
        D, stages, L, U, M, N, prec = {"c1": (2, 1, 2, 15, 1, 1024, "fp32"), "c2a": (8, 1, 2, 15, 1, 65536, "fp32"),
                                       "c5": (256, 8, 2, 256, 1, 1 << 20, "bf16")}[name]
        tnf.set_conditioner_precision(prec)
        nf = de.NormFlow(D, True, "coupling", stages, L, U)
        params = T(synthetic_params(chain_spec(nf.bijectors), D, M, seed=0)).cuda()

        def step():
            z, lq = nf.forward(params, N)
            return nf.log_prob(z, params)
        ms, launches, out = timed(step, steps=10 if name != "c5" else 5)
        n_cpl = 2 * stages
        flop = 2 * n_cpl * 4 * ((D // 2) * U + (L - 1) * U * U + U * (D // 2))
        byts = 2 * n_cpl * (2 * D * 4 + 8)
        rows = M * N
        line = {"workload": name, "config": "NormFlow(%d,conditioner,'coupling',%d,%d,%d) M=%d N=%d, %s" % (D, stages, L, U, M, N, prec),
                "value": rows / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "kernel_launches_per_step": launches,
                "roofline": {"tensor_frac": flop * rows / (ms * 1e-3) / 1e12 / PEAK["bf16_tflops"],
                             "hbm_frac_per_layer_bytes": byts * rows / (ms * 1e-3) / 1e9 / PEAK["hbm_gbs"],
                             "flop_per_sample": flop, "bytes_per_sample_per_layer_accounting": byts,
                             "bytes_per_sample_chain_fused": 2 * (2 * D * 4 + 4),
                             "hbm_frac_chain_fused_bytes": 2 * (2 * D * 4 + 4) * rows / (ms * 1e-3) / 1e9 / PEAK["hbm_gbs"]},
                "finite": bool(torch.isfinite(out).all())}
    else:   # conditional flows: one parameter row per sample (regime B)
        if name == "c2b":
            D, Dx, hidden, M, sup, U = 8, 8, [100], 65536, None, 15      # notebooks/LFI_gauss.ipynb:93-117
        else:   # c4like: scripts/lfi_mat.py shapes with 'coupling'
            D, Dx, hidden, M, U = 6, 2, [64, 64], 1 << 18, 15          # scripts/lfi_mat.py:35-46: 2 D clamped up to 15
            sup = ToInterval(D, [-2.0] * D, [2.0] * D)
        tnf.set_conditioner_precision("fp32")
        nf = de.NormFlow(D, True, "coupling", 1, 2, U, sup)
        torch.manual_seed(0)
        cde = ConditionalDensityEstimator(nf, Dx, hidden).cuda()
        x = torch.randn(M, Dx, device="cuda")
        with torch.no_grad():
            params = cde.param_net(x).contiguous()

        def step_flow():
            z, lq = nf.forward(params, 1)
            return nf.log_prob(z, params)

        def step_cde():
            z, lq = cde(x, N=1)
            return cde.log_prob(z, x)
        ms, launches, out = timed(step_flow)
        ms2, launches2, _ = timed(step_cde)
        # conditional log_prob alone (the SNPE density evaluation): hyper-network fused into the flow kernel vs params in HBM
        from torch_nf_b200 import config
        zz = (torch.rand(M, 1, D, device="cuda") * 3.6 - 1.8) if sup is not None else torch.randn(M, 1, D, device="cuda")
        config.set_cde_fusion(True)
        config.set_cde_variant("cc")
        ms_c, launches_c, lp_c = timed(lambda: cde.log_prob(zz, x))
        config.set_cde_variant("tc")
        ms_f, launches_f, lp_f = timed(lambda: cde.log_prob(zz, x))
        with torch.no_grad():
            hh = cde.param_net[:-1](x)
        ms_h, _, _ = timed(lambda: cde.param_net[:-1](x))
        config.set_cde_fusion(False)
        ms_u, launches_u, lp_u = timed(lambda: cde.log_prob(zz, x))
        config.set_cde_fusion(True)
        H = cde.param_net[-1].in_features
        pbytes = 2 * nf.D_params * 4            # the parameter row is read once per direction
        line = {"workload": name, "config": "conditional NormFlow D=%d, D_params=%d, M=%d, N=1, fp32 (per-sample weights)" % (D, nf.D_params, M),
                "value": M / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "kernel_launches_per_step": launches,
                "with_hyper_network": {"value": M / (ms2 * 1e-3), "ms_per_step": ms2, "note": "param_net (torch Linear/Tanh) evaluated inside the step, twice"},
                "cde_log_prob": {"fused": {"value": M / (ms_f * 1e-3), "ms": ms_f, "kernel_launches": launches_f,
                                           "producer": "tcgen05 (fp16 hi/lo split, 3 MMAs per product), parameter rows in tensor memory",
                                           "hidden_layers_of_param_net_ms": ms_h,
                                           "algorithmic_tflops": 2.0 * (H + 1) * nf.D_params * M / (ms_f * 1e-3) / 1e12,
                                           "hbm_bytes_per_sample": 4 * (H + D + 1)},
                                 "fused_cuda_cores": {"value": M / (ms_c * 1e-3), "ms": ms_c, "kernel_launches": launches_c,
                                                      "fp32_tflops": 2.0 * H * nf.D_params * M / (ms_c * 1e-3) / 1e12,
                                                      "max_rel_diff_vs_tc": float(((lp_f - lp_c).abs() / lp_c.abs().clamp(min=1)).max())},
                                 "unfused": {"value": M / (ms_u * 1e-3), "ms": ms_u, "kernel_launches": launches_u,
                                             "hbm_bytes_per_sample": 2 * nf.D_params * 4 + 4 * (H + D + 1)},
                                 "max_rel_diff": float(((lp_f - lp_u).abs() / lp_u.abs().clamp(min=1)).max())},
                "roofline": {"bound": "hbm (parameter rows)", "bytes_per_sample": pbytes + 2 * (2 * D * 4 + 4),
                             "hbm_frac": (pbytes + 2 * (2 * D * 4 + 4)) * M / (ms * 1e-3) / 1e9 / PEAK["hbm_gbs"]},
                "finite": bool(torch.isfinite(out).all())}
    print(json.dumps(line), flush=True)


for nm in (sys.argv[1:] or ["c1", "c2a", "c2b", "c4like", "c5"]):
    run(nm)
