"""GPU tests of the tcgen05 tensor-core coupling kernel (bf16 conditioner).

Tolerances (stated): against the fp32 reference/oracle
max|dz| <= 5e-2 and |d log_det| <= 2e-3*max(1,|.|)*layers-ish; against the
bf16-emulating oracle (same operand rounding, exact tanh) the kernel must
agree to ~1e-2, which separates rounding noise from layout bugs."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import flow_oracle as O
import torch_nf_b200 as tnf
import torch_nf_b200.density_estimator as de
from torch_nf_b200 import _lib, ops
from torch_nf_b200.synthetic import chain_spec, synthetic_params, synthetic_noise

pytestmark = pytest.mark.gpu
T = torch.tensor


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("a_in_tmem", [0, 1])
@pytest.mark.parametrize("K,N", [(16, 16), (32, 256), (32, 32), (64, 128), (256, 256), (256, 32), (128, 256), (256, 64)])
def test_umma_selftest_gemm(K, N, a_in_tmem):
    """One UMMA GEMM (A from SMEM image or from TMEM) through the kernel's own packing, descriptors and TMEM layouts."""
    g = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=g)
    W = torch.randn(K, N, generator=g) / np.sqrt(K)
    out = torch.zeros(128, N, device="cuda")
    Ad, Wd = A.cuda(), W.cuda()
    rc = _lib.lib().tnf_tc_selftest_gemm(Ad.data_ptr(), Wd.data_ptr(), out.data_ptr(), K, N, a_in_tmem,
                                         torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "tnf_tc_selftest_gemm")
    torch.cuda.synchronize()
    ref = _bf16(A).double() @ _bf16(W).double()
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 1e-4, "UMMA GEMM mismatch %g" % err


@pytest.mark.parametrize("D,U,L,N", [(64, 256, 2, 128), (64, 256, 2, 1000), (64, 256, 1, 300), (64, 128, 2, 257),
                                     (64, 64, 3, 129), (128, 256, 2, 384), (256, 256, 2, 200), (64, 256, 5, 140)])
@pytest.mark.parametrize("upper", [True, False])
def test_coupling_tc_vs_oracle(D, U, L, N, upper):
    assert ops.tc_supported(D, U, L)
    params = T(synthetic_params([("RealNVP", L, U, upper)], D, 1, seed=D + U + L))
    z_in = T(synthetic_noise(1, N, D, seed=5).astype(np.float32))
    packed = ops.tc_pack(params.cuda()[0], D, U, L, upper)
    for inverse in (False, True):
        direction = ops.TNF_INVERSE if inverse else ops.TNF_FORWARD
        z, ld = ops.coupling_tc(z_in.cuda(), packed, D, U, L, upper, direction)
        torch.cuda.synchronize()
        z, ld = z.cpu(), ld.cpu().view(1, N)
        ze, lde = O.coupling_bf16_emulated(z_in, params, D, L, U, upper, inverse)
        zr, ldr = (O.coupling_inverse if inverse else O.coupling_forward)(z_in, params, D, L, U, upper)
        h = D // 2
        keep = slice(0, h) if upper else slice(h, D)
        assert torch.equal(z[:, :, keep], z_in[:, :, keep])          # pass-through half bit-identical
        assert (z - ze).abs().max().item() < 2e-2 * L, "vs bf16-emulated oracle"
        assert (ld - lde).abs().max().item() < 2e-2 * L
        from torch_nf_b200 import config
        assert (z - zr).abs().max().item() <= config.BF16_TOL_Z, "vs fp32 oracle (stated bf16 tolerance)"
        assert (ld - ldr).abs().max().item() <= config.BF16_TOL_LD


@pytest.mark.parametrize("D,U,L,N", [(64, 256, 2, 1000), (64, 128, 2, 257), (128, 256, 2, 384), (64, 64, 3, 129)])
def test_coupling_tc_kernel_variants(D, U, L, N):
    """Diagnostic variants: 1 routes D <= 128 through the first (8 epilogue warp) kernel, 2 through the two-tile kernel
    without CTA pairs, 4 through the two-tile kernel on CTA pairs (tcgen05 cta_group::2), 3 through its N-half /
    TMEM-fed form (coupling_tc5_kernel, where the shape allows) without the FMA-pipe tanh share; 0 is the default
    (coupling_tc5_kernel with 1 of 8 tanh on the FMA pipe at the C3 shape).  Variants 2, 3 and 4 do the same
    arithmetic and must agree bit for bit; variant 1 differs by the bf16 hi/lo rounding of the bias MMA; the default
    by the polynomial tanh (max abs error 1.4e-3 on 1/8 of the activations)."""
    params = T(synthetic_params([("RealNVP", L, U, True)], D, 1, seed=3))
    z_in = T(synthetic_noise(1, N, D, seed=8).astype(np.float32)).cuda()
    packed = ops.tc_pack(params.cuda()[0], D, U, L, True)
    out = {}
    for variant in (0, 1, 2, 3, 4):
        for direction in (ops.TNF_FORWARD, ops.TNF_INVERSE):
            z, ld = ops.coupling_tc(z_in, packed, D, U, L, True, direction, variant=variant)
            torch.cuda.synchronize()
            out[(variant, direction)] = (z.cpu(), ld.cpu())
    for direction in (ops.TNF_FORWARD, ops.TNF_INVERSE):
        z0, l0 = out[(4, direction)]
        z1, l1 = out[(1, direction)]
        z2, l2 = out[(2, direction)]
        z3, l3 = out[(3, direction)]
        zd, ldd = out[(0, direction)]
        assert torch.equal(z0, z2) and torch.equal(l0, l2)
        assert torch.equal(z0, z3) and torch.equal(l0, l3)
        assert (z0 - zd).abs().max().item() < 2e-2 * L and (l0 - ldd).abs().max().item() < 2e-2 * L
        assert (z0 - z1).abs().max().item() < 2e-2 * L
        assert (l0 - l1).abs().max().item() < 2e-2 * L
        ze, lde = O.coupling_bf16_emulated(z_in.cpu(), params, D, L, U, True, direction == ops.TNF_INVERSE)
        assert (z1 - ze).abs().max().item() < 2e-2 * L
        assert (l1.view(1, N) - lde).abs().max().item() < 2e-2 * L


def test_coupling_tc_accum_and_preaffine():
    D, U, L, N = 64, 256, 2, 515
    params = T(synthetic_params([("RealNVP", L, U, True)], D, 1, seed=9))
    z_in = T(synthetic_noise(1, N, D, seed=6).astype(np.float32))
    packed = ops.tc_pack(params.cuda()[0], D, U, L, True)
    ps = (torch.rand(D) + 0.5).cuda()
    pb = torch.randn(D).cuda()
    base = torch.randn(N).cuda()
    for accum, sign in ((ops.TNF_LD_ADD, 1.0), (ops.TNF_LD_SUB, -1.0)):
        ld = base.clone()
        z, _ = ops.coupling_tc(z_in.cuda(), packed, D, U, L, True, ops.TNF_FORWARD, ld=ld, accum=accum,
                               pre_scale=ps, pre_shift=pb)
        zin2 = z_in * ps.cpu() + pb.cpu()
        ze, lde = O.coupling_bf16_emulated(zin2, params, D, L, U, True, False)
        assert (z.cpu() - ze).abs().max().item() < 4e-2
        assert (ld.cpu() - (base.cpu() + sign * lde.view(-1))).abs().max().item() < 4e-2


@pytest.mark.parametrize("D,U,L,N,upper", [(64, 256, 2, 1000, True), (64, 256, 2, 333, False), (64, 128, 2, 257, True),
                                            (128, 256, 2, 384, True), (64, 256, 1, 300, True), (64, 256, 3, 500, False),
                                            (128, 128, 2, 129, False), (64, 256, 5, 140, True)])
def test_coupling_fp32_parity_on_tensor_cores(D, U, L, N, upper):
    """TNF_TC_FP32 (fp16 hi/lo operand split, three MMAs per product, fp32-accurate tanh / exp) against the fp32
    oracle with the FP32 tolerance of the north star: max|dz|/max(1,|z|) <= 1e-5, log-det <= 1e-4 relative."""
    from torch_nf_b200 import config
    params = T(synthetic_params([("RealNVP", L, U, upper)], D, 1, seed=3))
    z = torch.randn(1, N, D, generator=torch.Generator().manual_seed(1)) * 1.3
    packed = ops.tc_pack(params.cuda()[0], D, U, L, upper, precision="fp32_tc")
    ps = (torch.rand(D) + 0.5); pb = torch.randn(D) * 0.3
    for direction, fn in ((ops.TNF_FORWARD, O.coupling_forward), (ops.TNF_INVERSE, O.coupling_inverse)):
        zo, ldo = fn(z, params, D, L, U, upper)
        before = _lib.launch_count()
        zd, ld = ops.coupling_tc(z.cuda(), packed, D, U, L, upper, direction, precision="fp32_tc")
        assert _lib.launch_count() == before + 1
        h = D // 2
        keep = slice(0, h) if upper else slice(h, D)
        assert torch.equal(zd.cpu()[:, :, keep], z[:, :, keep])          # pass-through half bit-identical
        rz = ((zd.cpu() - zo).abs() / zo.abs().clamp(min=1)).max().item()
        rl = ((ld.cpu().view(1, N) - ldo).abs() / ldo.abs().clamp(min=1)).max().item()
        print("fp32_tc D=%d U=%d L=%d dir=%d: rel z %.3g, rel log-det %.3g" % (D, U, L, direction, rz, rl))
        assert rz <= config.FP32_TOL_Z and rl <= config.FP32_TOL_LOGP, (rz, rl)
        # folded pre-affine and log-det accumulation
        base = torch.randn(N)
        ld2 = base.clone().cuda()
        z2, _ = ops.coupling_tc(z.cuda(), packed, D, U, L, upper, direction, ld=ld2, accum=ops.TNF_LD_SUB,
                                pre_scale=ps.cuda(), pre_shift=pb.cuda(), precision="fp32_tc")
        zo2, ldo2 = fn(z * ps + pb, params, D, L, U, upper)
        assert ((z2.cpu() - zo2).abs() / zo2.abs().clamp(min=1)).max().item() <= config.FP32_TOL_Z
        assert ((ld2.cpu() - (base - ldo2.view(-1))).abs()).max().item() <= 1e-4 * max(1.0, float(ldo2.abs().max()))


def _bf16_chain_vs_golden(golden, name):
    """Whole chain in bf16-conditioner mode vs the reference's golden vectors with THE stated bf16 tolerance
    (torch_nf_b200.config.BF16_TOL_Z / BF16_TOL_LOGP)."""
    from torch_nf_b200 import config
    g = golden(name)
    D, stages, L, U, M, N, pseed, oseed = [int(v) for v in g["cfg"]]
    nf = de.NormFlow(D, True, "coupling", stages, L, U)
    params = T(synthetic_params(chain_spec(nf.bijectors), D, M, seed=pseed)).cuda()
    np.random.seed(oseed)
    omega = np.random.normal(0.0, 1.0, (M, N, D))
    tnf.set_conditioner_precision("bf16")
    old_min = config.tc_min_rows()
    config.set_tc_min_rows(1)           # the goldens are small (C5: 64 rows): still the tensor-core kernels
    try:
        before = _lib.launch_count()
        with torch.no_grad():
            z, lq = nf.forward(params, N, omega=omega)
            lp = nf.log_prob(T(g["z"]).cuda(), params)
        assert _lib.launch_count() - before > 2 * stages * 2
    finally:
        tnf.set_conditioner_precision("fp32")
        config.set_tc_min_rows(old_min)
    dz = np.abs(z.cpu().numpy() - g["z"]).max()
    dlq = (np.abs(lq.cpu().numpy() - g["log_q_z"]) / np.maximum(1.0, np.abs(g["log_q_z"]))).max()
    dlp = (np.abs(lp.cpu().numpy() - g["log_prob"]) / np.maximum(1.0, np.abs(g["log_prob"]))).max()
    print("%s bf16 chain: max|dz| = %.3g, rel log_q = %.3g, rel log_prob = %.3g" % (name, dz, dlq, dlp))
    assert dz > 1e-5, "the bf16 path did not run"
    assert dz <= config.BF16_TOL_Z, dz
    assert dlq <= config.BF16_TOL_LOGP and dlp <= config.BF16_TOL_LOGP, (dlq, dlp)


def test_normflow_bf16_mode_c3(golden):
    _bf16_chain_vs_golden(golden, "flow_c3")


def test_normflow_bf16_mode_c5(golden):
    """BASELINE.json configuration 5 (D = 256, 16 coupling layers) in ITS precision, as a chain."""
    _bf16_chain_vs_golden(golden, "flow_c5")


@pytest.mark.parametrize("D,stages,N", [(64, 4, 1 << 16), (256, 8, 1 << 13)])
def test_normflow_bf16_mode_vs_oracle_at_scale(D, stages, N):
    """C3 over 2^16 rows and C5 over 2^13 rows against the CPU oracle (the golden vectors hold 192 / 64 rows)."""
    from torch_nf_b200 import config
    L, U = 2, 256
    chain = O.build_chain(D, "coupling", stages, L, U)
    nf = de.NormFlow(D, True, "coupling", stages, L, U)
    params = T(synthetic_params(chain_spec(nf.bijectors), D, 1, seed=0))
    omega = np.random.RandomState(5).standard_normal((1, N, D))
    with torch.no_grad():
        zo, lqo, st = O.normflow_forward(chain, D, params, omega)
        lpo = O.normflow_log_prob(chain, D, zo, params, st)
    tnf.set_conditioner_precision("bf16")
    try:
        with torch.no_grad():
            z, lq = nf.forward(params.cuda(), N, omega=omega)
            lp = nf.log_prob(zo.cuda(), params.cuda())
    finally:
        tnf.set_conditioner_precision("fp32")
    dz = (z.cpu() - zo).abs().max().item()
    dlq = ((lq.cpu() - lqo).abs() / lqo.abs().clamp(min=1.0)).max().item()
    dlp = ((lp.cpu() - lpo).abs() / lpo.abs().clamp(min=1.0)).max().item()
    print("D=%d, %d layers, %d rows, bf16: max|dz| = %.3g, rel log_q = %.3g, rel log_prob = %.3g" % (D, 2 * stages, N, dz, dlq, dlp))
    assert 1e-5 < dz <= config.BF16_TOL_Z, dz
    assert dlq <= config.BF16_TOL_LOGP and dlp <= config.BF16_TOL_LOGP, (dlq, dlp)


def test_coupling_tc_fused_column_stats():
    """The kernel's own [sum | sumsq | rows] of its output equals tnf_colstats on that output: the two-tile bf16 kernels
    (fp32 partial sums), and the TMEM-resident kernel in both precisions (float64 accumulation: fp32-parity class)."""
    U, L = 256, 2
    cases = [(64, 1000, True, "bf16"), (64, 128 * 148 * 2 + 77, False, "bf16"), (64, 90, True, "bf16"), (128, 128 * 148 + 5, True, "bf16"),
             (64, 128 * 148 * 3 + 77, False, "fp32_tc"), (64, 90, True, "fp32_tc"), (128, 128 * 148 + 5, True, "fp32_tc"),
             (256, 128 * 148 * 2 + 9, True, "bf16"), (256, 300, False, "fp32_tc")]
    for D, N, upper, prec in cases:
        params = T(synthetic_params([("RealNVP", L, U, upper)], D, 1, seed=21))
        z_in = torch.randn(1, N, D, device="cuda") * 1.5 + 0.3
        packed = ops.tc_pack(params.cuda()[0], D, U, L, upper, precision=prec)
        ps = (torch.rand(D) + 0.5).cuda(); pb = torch.randn(D).cuda()
        z, ld, sums = ops.coupling_tc(z_in, packed, D, U, L, upper, ops.TNF_FORWARD, pre_scale=ps, pre_shift=pb,
                                      want_stats=True, precision=prec)
        ref = ops.colstats(z, D)
        assert float(sums[2 * D]) == N == float(ref[2 * D])
        tc6 = prec == "fp32_tc" or D == 256
        np.testing.assert_allclose(sums[:2 * D].cpu().numpy(), ref[:2 * D].cpu().numpy(), rtol=2e-6 if tc6 else 2e-5,
                                   atol=(1e-4 if tc6 else 2e-3) * N ** 0.5)
        z2, _ = ops.coupling_tc(z_in, packed, D, U, L, upper, ops.TNF_FORWARD, pre_scale=ps, pre_shift=pb, precision=prec)
        assert torch.equal(z, z2)


def test_cta_pair_kernel_matches_single_cta_kernel_at_scale():
    """The CTA-pair kernel (cross-CTA mbarrier arrivals, multicast commits, forwarded stage barriers) against the
    single-CTA two-tile kernel on many tiles per CTA and repeated launches: any lost ordering between the two CTAs of
    a pair would show up as a bit difference somewhere in 2^18 rows."""
    D, U, L, N = 64, 256, 2, (1 << 18) + 333
    params = T(synthetic_params([("RealNVP", L, U, False)], D, 1, seed=12))
    packed = ops.tc_pack(params.cuda()[0], D, U, L, False)
    z_in = torch.randn(1, N, D, device="cuda")
    ps = (torch.rand(D) + 0.5).cuda(); pb = torch.randn(D).cuda()
    ref = {}
    for direction in (ops.TNF_FORWARD, ops.TNF_INVERSE):
        z, ld, sums = ops.coupling_tc(z_in, packed, D, U, L, False, direction, pre_scale=ps, pre_shift=pb, want_stats=True,
                                      variant=2)
        ref[direction] = (z.clone(), ld.clone(), sums.clone())
    for rep in range(3):
        for direction in (ops.TNF_FORWARD, ops.TNF_INVERSE):
            for variant in (3, 4):      # both CTA-pair kernels (N-half / TMEM-fed and plain)
                z, ld, sums = ops.coupling_tc(z_in, packed, D, U, L, False, direction, pre_scale=ps, pre_shift=pb,
                                              want_stats=True, variant=variant)
                assert torch.equal(z, ref[direction][0]) and torch.equal(ld, ref[direction][1])
            np.testing.assert_allclose(sums.cpu().numpy(), ref[direction][2].cpu().numpy(), rtol=2e-5, atol=2e-3 * N ** 0.5)
    assert torch.isfinite(z).all()
