#!/usr/bin/env python
"""Headline benchmark: flow log_prob+sample samples/sec at D=64, 8 coupling layers.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch: ``NormFlow.forward``
(draw base noise, run the chain, accumulate the log-density) followed by
``NormFlow.log_prob`` on the samples just produced, for the configuration
BASELINE.json quotes the metric on: ``NormFlow(64, False, 'coupling', 4, 2, 256)``
(C3: D=64, 8 RealNVP layers with 256-wide 2-layer conditioners, 8 BatchNorm, 4
Affine), batch 2^20 per GPU (weak scaling; BatchNorm statistics are all-reduced
so every rank normalises with the global batch).  Prints ONE JSON line.

`value`   device-resident throughput (weights and samples stay in HBM),
          conditioner GEMMs in bf16 on tcgen05 (dtype "bf16"; affine transform,
          log-det and BatchNorm in fp32).  The same step at the REFERENCE'S
          precision (fp32 parity on tensor cores, fp16 hi/lo operand split) is
          measured in the same run and reported in the `fp32` object.
`e2e`     the same step through the public API with HOST tensors (pinned):
          parameters and samples cross PCIe inside the timed region.
`--workload train`  BASELINE config 4 (conditional-flow training step), `--workload train_c3`  the C3 flow's
maximum-likelihood training step (tensor-core backward), `--config c5`  BASELINE config 5.
`--impl reference`  times the reference's CPU algorithm (the oracle port, torch
          CPU ops on all host cores) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, STAGES, L, U = 64, 4, 2, 256
N_LAYERS = 2 * STAGES
BATCH_PER_GPU = 1 << 20
FLOP_PER_SAMPLE_LAYER = 4 * (32 * U + (L - 1) * U * U + U * 32)       # 327 680
BYTES_PER_SAMPLE_LAYER = 2 * D * 4 + 8                                 # 520
METRIC = "flow log_prob+sample samples/sec at D=64,L=8"
WORKLOAD = "C3: NormFlow(64,False,'coupling',4,2,256) sample(N)+log_prob, batch 2^20 per GPU"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(source="measured", tflops_burst=p["bf16_tflops"], tflops_sustained=p["bf16_tflops_sustained"],
                    hbm_gbs=p["hbm_gbs"])
    return dict(source="fallback", tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0)


class ClockSampler(object):
    """SM clock, power and throttle reasons sampled DURING the timed region: NVML polled every 5 ms from a thread
    (nvidia-smi -lms as the fallback when the NVML bindings are missing)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason* bits
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.samples, self.stop_flag = None, [], False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES remaps indices: address the device by the UUID torch reports when possible
            handle = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(torch.cuda.current_device()).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = (pynvml, handle)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv, h = self.nvml
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = 0.0
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append((sm, mx, pw, rs))
            except Exception:
                pass
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        sm, mx, pw, reasons = [], [], [], set()
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            for s_, m_, p_, r_ in self.samples:
                sm.append(s_); mx.append(m_); pw.append(p_)
                for nm, bit in self.BITS.items():
                    if r_ & bit:
                        reasons.add(nm)
        else:
            if self.proc is None:
                return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for ln in self.lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        # under load = the upper half of the samples (the sampler also sees the idle edges)
        hot = sorted(sm)[len(sm) // 2:]
        return dict(sm_mhz=statistics.median(hot), sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm),
                    reasons=sorted(reasons))


def cpu_port_throughput(n_rows, repeats, threads):
    """The reference's CPU algorithm (oracle port) on `n_rows` samples of the same
    workload: one forward (incl. the reference's host numpy draw and float64 base
    density) + one log_prob.  Returns (best samples/s, seconds of the best pass)."""
    import numpy as np
    import torch
    from oracle import flow_oracle as O
    from torch_nf_b200.synthetic import synthetic_params
    torch.set_num_threads(threads)
    chain = O.build_chain(D, "coupling", STAGES, L, U)
    spec = [(b["kind"], b.get("L", 0), b.get("U", 0), b.get("upper", False)) for b in chain]
    params = torch.tensor(synthetic_params(spec, D, 1, seed=0))
    best = None
    with torch.no_grad():
        for i in range(repeats + 1):            # first pass is the warm-up
            t0 = time.perf_counter()
            omega = np.random.normal(0.0, 1.0, (1, n_rows, D))     # density_estimator.py:366
            z, lq, st = O.normflow_forward(chain, D, params, omega)
            lp = O.normflow_log_prob(chain, D, z, params, st)
            dt = time.perf_counter() - t0
            if i > 0 and (best is None or dt < best):
                best = dt
    assert torch.isfinite(lp).all()
    return n_rows / best, best


def run_reference(args):
    """Reference arm: rank 0 only, host cores only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    n_rows = 1 << 14
    import numpy as np
    from oracle import flow_oracle as O
    from torch_nf_b200.synthetic import synthetic_params
    torch.set_num_threads(threads)
    chain = O.build_chain(D, "coupling", STAGES, L, U)
    spec = [(b["kind"], b.get("L", 0), b.get("U", 0), b.get("upper", False)) for b in chain]
    params = torch.tensor(synthetic_params(spec, D, 1, seed=0))

    def step():
        omega = np.random.normal(0.0, 1.0, (1, n_rows, D))
        z, lq, st = O.normflow_forward(chain, D, params, omega)
        return O.normflow_log_prob(chain, D, z, params, st)

    with torch.no_grad():
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
    value = n_rows * args.steps / dt
    sample = "%d of 2^20 rows per step, torch CPU ops, %d threads" % (n_rows, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_sample_rows_per_step": n_rows, "same_config": False,
                   "note": "bounded sample: 2^14 of the 2^20 rows per step, same flow and weights"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample,
                         "same_config": False, "rows_per_step": n_rows},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="samples per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "fp32_cc"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch rows per GPU (default); strong: --batch rows in total, split over the ranks")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32-parity measurement of the same step")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step measurement of the same flow")
    ap.add_argument("--config", default="c3", choices=["c3", "c5"],
                    help="c3 (default): the configuration the metric is quoted on; c5: BASELINE.json config 5 "
                         "(D=256, 16 coupling layers, bf16 conditioner; quoted at batch 2^20 over 8 GPUs: --scaling strong)")
    ap.add_argument("--workload", default="sample_logprob", choices=["sample_logprob", "train", "train_c3"],
                    help="train: BASELINE.json config 4, the conditional-flow training step (forward + backward + gradient "
                         "all-reduce + Adam) of profiles/scripts/bench_train.py; train_c3: the maximum-likelihood training "
                         "step of the C3 flow with the tensor-core backward (profiles/scripts/bench_train_c3.py); both print "
                         "that script's JSON line")
    args = ap.parse_args()
    if args.workload in ("train", "train_c3"):
        import runpy
        script = "bench_train.py" if args.workload == "train" else "bench_train_c3.py"
        sys.argv = [os.path.join(ROOT, "profiles", "scripts", script), "--steps", str(args.steps), "--warmup", str(max(args.warmup, 3))]
        runpy.run_path(sys.argv[0], run_name="__main__")
        return
    if args.config == "c5":
        global D, STAGES, N_LAYERS, FLOP_PER_SAMPLE_LAYER, BYTES_PER_SAMPLE_LAYER, METRIC, WORKLOAD
        D, STAGES = 256, 8
        N_LAYERS = 2 * STAGES
        FLOP_PER_SAMPLE_LAYER = 4 * (128 * U + (L - 1) * U * U + U * 128)      # 524 288
        BYTES_PER_SAMPLE_LAYER = 2 * D * 4 + 8                                  # 2 056
        METRIC = "flow log_prob+sample samples/sec at D=256,L=16 (BASELINE config 5)"
        WORKLOAD = "C5: NormFlow(256,False,'coupling',8,2,256) sample(N)+log_prob, bf16 conditioner / fp32 log-det"
        args.no_fp32 = True
        args.no_cpu_baseline = True
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as td
    import torch_nf_b200 as tnf
    import torch_nf_b200.density_estimator as de
    from torch_nf_b200 import _lib, dist, ops
    from torch_nf_b200.synthetic import chain_spec, synthetic_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
        dist.enable()
        # BatchNorm statistics across ranks: inside the fold kernels over NVLink peer memory (falls back to one NCCL
        # all-reduce per BatchNorm when symmetric memory is unavailable; TNF_PEER_EXCHANGE=0 forces the fallback)
        if os.environ.get("TNF_PEER_EXCHANGE", "1") != "0":
            dist.enable_peer_exchange()
    tnf.set_conditioner_precision(args.precision)
    _lib.lib()

    B = args.batch if args.scaling == "weak" else -(-args.batch // world)
    nf = de.NormFlow(D, True, "coupling", STAGES, L, U)
    params_host = torch.tensor(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=0)).pin_memory()
    params = params_host.to(dev)
    np.random.seed(1234 + rank)          # per-rank Philox sub-stream

    def step_resident():
        z, lq = nf.forward(params, B)
        lp = nf.log_prob(z, params)
        return z, lq, lp

    def step_host():
        z, lq = nf.forward(params_host, B)            # CPU tensors in, CPU (pinned) tensors out
        lp = nf.log_prob(z, params_host)
        return z, lq, lp

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, timer=None):
        with torch.no_grad():
            for _ in range(warmup):
                out = fn()
            barrier()
            if timer is not None:
                ops.kernel_timer = timer
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = _lib.launch_count()
            e0.record()
            for _ in range(steps):
                out = fn()
            e1.record()
            barrier()
            ops.kernel_timer = None
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            td.all_reduce(t, op=td.ReduceOp.MAX)
            ms = float(t.item())
        return ms, _lib.launch_count() - l0, out

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    timer = []
    ms, launches, out = timed(step_resident, args.steps, args.warmup, timer)
    clock_info = clocks.stop() if rank == 0 else None
    z, lq, lp = out
    finite = bool(torch.isfinite(lp).all().item())
    # size-independent property at full size: log_prob(sample) == the sampler's own log-density
    consistency = float((lq.float() - lp).abs().max().item())
    value = world * B * args.steps / (ms * 1e-3)

    # dominant kernel: CUDA events recorded around every tensor-core coupling launch of the timed region
    kern_ms = [a.elapsed_time(b) for (a, b) in timer]
    pk = peaks()

    def roofline_of(kern_ms, step_ms, steps, kernel, split_factor):
        if not kern_ms:
            return None
        avg = sum(kern_ms) / len(kern_ms)
        achieved = FLOP_PER_SAMPLE_LAYER * B / (avg * 1e-3) / 1e12
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "tc_traffic.json")
        if os.path.exists(tpath) and args.config == "c3":       # dram__bytes_read+write per launch from the committed `ncu --set full` capture,
            with open(tpath) as fh:     # valid only for the kernel source it was taken on (digest checked here)
                tj = json.load(fh)
            ent = tj.get(kernel)
            if ent:
                from torch_nf_b200 import _build
                import hashlib
                with open(os.path.join(_build.CSRC, ent["source"]), "rb") as fh:
                    dig = hashlib.sha256(fh.read()).hexdigest()[:16]
                traffic = ent["dram_bytes_per_launch"]
                traffic_src = "ncu capture %s (%s)" % (ent["capture"], "current source" if dig == ent["source_sha256_16"]
                                                      else "STALE: kernel source changed since the capture")
        return {"bound": "tensor", "achieved": achieved, "peak": pk["tflops_burst"], "unit": "TFLOP/s",
                "frac": achieved / pk["tflops_burst"], "traffic": traffic, "traffic_source": traffic_src,
                "kernel": kernel, "launches_timed": len(kern_ms), "avg_launch_ms": avg,
                "peak_source": pk["source"] + " (burst bf16 GEMM: the timed region is ~0.1 s at full clocks; sustained %.1f -> frac %.3f)" % (
                    pk["tflops_sustained"], achieved / pk["tflops_sustained"]),
                "algorithmic_flop_per_launch": FLOP_PER_SAMPLE_LAYER * B, "mma_passes_per_product": split_factor,
                "kernel_share_of_step": sum(kern_ms) / (step_ms * steps),
                "chain_roofline_frac": (world * B * steps / (step_ms * steps * 1e-3)) / world / (
                    1.0 / (2 * N_LAYERS * FLOP_PER_SAMPLE_LAYER / (pk["tflops_burst"] * 1e12))),
                "hbm_gbs_at_algorithmic_bytes": BYTES_PER_SAMPLE_LAYER * B / (avg * 1e-3) / 1e9,
                "hbm_frac_at_algorithmic_bytes": BYTES_PER_SAMPLE_LAYER * B / (avg * 1e-3) / 1e9 / pk["hbm_gbs"]}

    roofline = roofline_of(kern_ms, ms / args.steps, args.steps,
                           "coupling_tc5_kernel" if (args.precision == "bf16" and args.config == "c3") else "coupling_tc6_kernel",
                           1 if args.precision == "bf16" else 3)

    # the same step at the reference's precision (fp32 parity mode), device resident, a few steps
    fp32 = None
    if args.precision == "bf16" and not args.no_fp32:
        tnf.set_conditioner_precision("fp32")
        t32 = []
        f_steps = max(2, min(args.steps, 4))
        ms32, _, out32 = timed(step_resident, f_steps, 3, t32)
        tnf.set_conditioner_precision(args.precision)
        k32 = [a.elapsed_time(b) for (a, b) in t32]
        z32, lq32, lp32 = out32
        fp32 = {"value": world * B * f_steps / (ms32 * 1e-3), "unit": "samples/s", "ms_per_step": ms32 / f_steps, "steps": f_steps,
                "dtype": "fp32 parity: fp16 hi/lo operand split on tcgen05, fp32 accumulate, rel 1e-5 on samples",
                "max_abs_logq_minus_logprob": float((lq32.float() - lp32).abs().max().item()),
                "roofline": roofline_of(k32, ms32 / f_steps, f_steps, "coupling_tc6_kernel", 3)}
        del out32, z32, lq32, lp32

    e2e = None
    if not args.no_e2e:
        e_steps = max(2, min(args.steps, 5))
        ms_h, _, out_h = timed(step_host, e_steps, 3)   # 3 warm-ups: both alternating pinned return buffers exist
        zh, lqh, lph = out_h
        assert not zh.is_cuda and not lph.is_cuda
        h2d = 2 * params_host.numel() * 4 + zh.numel() * 4
        d2h = zh.numel() * 4 + lqh.numel() * 8 + lph.numel() * 4
        e2e = {"value": world * B * e_steps / (ms_h * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_h / e_steps}

    # the same flow's maximum-likelihood TRAINING step (forward + tensor-core backward + Adam) on this GPU, as an extra
    # object of the line (1 GPU, bf16 mode, C3 only); `--workload train_c3` is the multi-GPU form of this measurement
    train_c3 = None
    if world == 1 and args.config == "c3" and args.precision == "bf16" and not args.no_train:
        try:
            del out
            torch.cuda.empty_cache()
            gen = torch.Generator(device=dev).manual_seed(7)
            z_data = (torch.randn(1, B, D, device=dev, generator=gen) * 1.2 + 0.1).contiguous()
            p_train = params.clone().requires_grad_(True)
            opt = torch.optim.Adam([p_train], lr=1e-4)
            t_losses = []

            def train_step_fn():
                opt.zero_grad(set_to_none=True)
                loss = -nf.log_prob(z_data, p_train).mean()
                loss.backward()
                opt.step()
                t_losses.append(loss.detach())
            for _ in range(2):
                train_step_fn()
            torch.cuda.synchronize()
            t_steps = 3
            bt = []
            ops.bwd_kernel_timer = bt
            l0 = _lib.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(t_steps):
                train_step_fn()
            e1.record()
            torch.cuda.synchronize()
            ops.bwd_kernel_timer = None
            ms_t = e0.elapsed_time(e1) / t_steps
            k_ms = [a.elapsed_time(b) for (a, b, _r) in bt]
            k_rows = sum(r for (_a, _b, r) in bt)
            pk = peaks()
            k_tflops = 2 * FLOP_PER_SAMPLE_LAYER * k_rows / (sum(k_ms) * 1e-3) / 1e12 if k_ms else None
            t_losses = [float(v) for v in t_losses]
            train_c3 = {"metric": "maximum-likelihood training step (-mean log_prob: forward + backward of %d coupling layers on tensor cores, "
                                  "BatchNorm / Affine backward, Adam)" % N_LAYERS,
                        "value": B / (ms_t * 1e-3), "unit": "samples/s", "ms_per_step": ms_t, "steps": t_steps, "rows": B,
                        "gpu_launches_per_step": int(_lib.launch_count() - l0) // t_steps,
                        "tflops_algorithmic": N_LAYERS * 4 * FLOP_PER_SAMPLE_LAYER * B / (ms_t * 1e-3) / 1e12,
                        "roofline": {"bound": "tensor", "kernel": "coupling_tcb_kernel", "launches_timed": len(k_ms),
                                     "avg_launch_ms": (sum(k_ms) / len(k_ms)) if k_ms else None,
                                     "achieved": k_tflops, "peak": pk["tflops_burst"], "unit": "TFLOP/s",
                                     "frac": (k_tflops / pk["tflops_burst"]) if k_tflops else None,
                                     "algorithmic_flop_per_launch": "2 x %d x rows (conditioner recompute + data gradient; the weight gradient runs in library GEMMs)" % FLOP_PER_SAMPLE_LAYER,
                                     "kernel_share_of_step": (sum(k_ms) / t_steps / ms_t) if k_ms else None},
                        "loss_first": t_losses[0], "loss_last": t_losses[-1], "finite": bool(np.isfinite(t_losses).all())}
            del z_data, p_train, opt
        except Exception as exc:       # the extra object must never cost the headline line
            train_c3 = {"error": "%s: %s" % (type(exc).__name__, exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        v, secs = cpu_port_throughput(1 << 14, 3, threads)
        cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port", "same_config": False,
               "rows_per_step": 1 << 14, "gpu_rows_per_step": B,
               "sample": "2^14 of 2^20 rows per step (same flow, same weights; samples/s is per-row), 1 warm-up + best of 3 "
                         "(%.2f s/pass), oracle port on torch CPU ops" % secs}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B, "l2": "inputs_larger_than_l2 (z = %d MB per pass)" % (B * D * 4 // 1000000), "rows_per_gpu": B,
                       "parallelism": "dp%d over sample rows" % world, "weights": "fan-in scaled synthetic, seed 0",
                       "noise": "device Philox4x32-10",
                       "bn_statistics_exchange": ("none (1 rank)" if world == 1 else
                                                  ("in-kernel over NVLink peer memory" if dist.peer_struct(0) is not None
                                                   else "NCCL all-reduce per BatchNorm (%s)" % (dist.peer_error or "peer exchange off")))},
            "roofline": roofline, "fp32": fp32, "train_c3": train_c3, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches) * world,
            "clocks": clock_info,
            "checks": {"finite": finite, "max_abs_logq_minus_logprob": consistency},
        }
        print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
