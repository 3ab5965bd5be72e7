"""GPU parity tests of the bijector kernels (through the C ABI) against the CPU
oracle and the golden vectors of the unmodified reference.  The structure
follows the reference's tests/test_bijectors.py."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as O
from torch_nf_b200.bijectors import RealNVP, Affine, BatchNorm, ToInterval, ToSimplex

pytestmark = pytest.mark.gpu
T = torch.tensor


def close(a, b, rtol, atol):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def test_native_library_loaded():
    from torch_nf_b200 import _lib
    assert _lib.lib().tnf_abi_version() == 3
    before = _lib.launch_count()
    Affine(4).forward_and_log_det(torch.zeros(1, 2, 4), torch.zeros(1, 8))
    assert _lib.launch_count() > before


def test_RealNVP_reference_style():
    """tests/test_bijectors.py:41-123 with the CUDA path as the implementation."""
    D, num_layers, num_units = 4, 2, 15
    real_nvp = RealNVP(D, num_layers, num_units)
    D_theta = real_nvp.count_num_params()
    M, N = 10, 5
    np.random.seed(0)
    params = T(np.random.normal(0.0, 0.1, (M, D_theta)))            # float64, as the reference test
    z_in = T(np.random.normal(0.0, 1.0, (M, N, D)))
    z, log_det = real_nvp(z_in, params)
    assert z.shape == (M, N, D) and log_det.shape == (M, N)
    assert z.dtype == torch.float64 and z.device == z_in.device
    assert torch.eq(z[:, :, : D // 2], z_in[:, :, : D // 2]).all()  # pass-through half bit-identical
    assert not torch.eq(z[:, :, D // 2:], z_in[:, :, D // 2:]).all()
    z_inv, log_det_inv = real_nvp.inverse_and_log_det(z, params)
    assert np.sum((z_in.numpy() - z_inv.numpy()) ** 2) < 1e-20
    assert np.sum((log_det.numpy() - log_det_inv.numpy()) ** 2) < 1e-20

    real_nvp = RealNVP(D, num_layers, num_units, transform_upper=False)
    params = T(np.random.normal(0.0, 1.0, (M, D_theta + 10)))       # extra trailing params ignored
    z, log_det = real_nvp(z_in, params)
    assert not torch.eq(z[:, :, : D // 2], z_in[:, :, : D // 2]).all()
    assert torch.eq(z[:, :, D // 2:], z_in[:, :, D // 2:]).all()
    zo, ldo = O.coupling_forward(z_in, params, D, num_layers, num_units, False)
    close(z, zo, 1e-12, 1e-12); close(log_det, ldo, 1e-12, 1e-12)

    for D, M in ((5, 20), (8, 20)):                                  # odd D, D = 8
        real_nvp = RealNVP(D, 1, 15, transform_upper=False)
        params = T(np.random.normal(0.0, 0.1, (M, real_nvp.count_num_params())))
        z_in = T(np.random.normal(0.0, 1.0, (M, N, D)))
        z, log_det = real_nvp(z_in, params)
        z_inv, log_det_inv = real_nvp.inverse_and_log_det(z, params)
        assert np.sum((z_in.numpy() - z_inv.numpy()) ** 2) < 1e-4
        assert np.sum((log_det.numpy() - log_det_inv.numpy()) ** 2) < 1e-4


REALNVP_CASES = ["d4_up_f64", "d4_lo_f64", "d5_lo_f64", "d5_up_f32", "d8_lo_f32", "d8_up_a_f32", "d6_up_b_f32"]


@pytest.mark.parametrize("nm", REALNVP_CASES)
def test_RealNVP_golden(golden, nm):
    g = golden("realnvp")
    D, L, U, up, M, N = [int(v) for v in g[nm + "_cfg"]]
    b = RealNVP(D, L, U, transform_upper=bool(up))
    params, z_in = T(g[nm + "_params"]), T(g[nm + "_z_in"])
    tol = 1e-12 if params.dtype == torch.float64 else 3e-6
    z, ld = b(z_in, params)
    close(z, g[nm + "_z_fwd"], tol, tol); close(ld, g[nm + "_ld_fwd"], tol, tol)
    z, ld = b.inverse_and_log_det(z_in, params)
    close(z, g[nm + "_z_inv"], tol, tol); close(ld, g[nm + "_ld_inv"], tol, tol)


def test_RealNVP_headline_shape_fp32(golden):
    """D=64, U=256, L=2 shared weights on the exact path: fp32 tolerance of the
    north star (max|dz|/max(1,|z|) <= 1e-5, log-det 1e-4)."""
    from torch_nf_b200.synthetic import synthetic_params, synthetic_noise
    g = golden("realnvp")
    for nm, up in (("d64_up_f32", True), ("d64_lo_f32", False)):
        D, L, U, _, M, N = [int(v) for v in g[nm + "_cfg"]]
        ps, zs = [int(v) for v in g[nm + "_seeds"]]
        params = T(synthetic_params([("RealNVP", L, U, up)], D, M, seed=ps)).cuda()
        z_in = T(synthetic_noise(M, N, D, seed=zs).astype(np.float32)).cuda()
        b = RealNVP(D, L, U, transform_upper=up)
        for fn, tag in ((b.forward_and_log_det, "fwd"), (b.inverse_and_log_det, "inv")):
            z, ld = fn(z_in, params)
            assert z.is_cuda
            zr, ldr = g[nm + "_z_" + tag], g[nm + "_ld_" + tag]
            assert np.max(np.abs(z.cpu().numpy() - zr) / np.maximum(1.0, np.abs(zr))) <= 1e-5
            assert np.max(np.abs(ld.cpu().numpy() - ldr) / np.maximum(1.0, np.abs(ldr))) <= 1e-4


def test_Affine():
    """tests/test_bijectors.py:273-297 closed form."""
    D = 4
    affine = Affine(D)
    M, N = 20, 50
    D_theta = affine.count_num_params()
    params = T(np.random.normal(0.0, 1.0, (M, D_theta))).float()
    z_in = T(np.random.normal(0.0, 1.0, (M, N, D))).float()
    z, log_det = affine.forward_and_log_det(z_in, params)
    log_scale, shift = params[:, :D], params[:, D:]
    z_true = z_in * torch.exp(log_scale[:, None, :]) + shift[:, None, :]
    log_det_true = torch.sum(log_scale, dim=1, keepdim=True)
    assert log_det.shape == (M, 1)
    assert np.sum((z.numpy() - z_true.numpy()) ** 2) < 1e-10
    assert np.sum((log_det.numpy() - log_det_true.numpy()) ** 2) < 1e-10
    z_inv, log_det_inv = affine.inverse_and_log_det(z, params)
    assert np.sum((z_in.numpy() - z_inv.numpy()) ** 2) < 1e-10
    assert np.sum((log_det.numpy() - log_det_inv.numpy()) ** 2) < 1e-10
    # shared weights (one parameter row), large N, resident on the GPU
    z_in = torch.randn(1, 5000, D, device="cuda")
    z, ld = affine(z_in, params[:1].cuda())
    zo, ldo = O.affine_forward(z_in.cpu(), params[:1], D)
    close(z, zo, 1e-6, 1e-6); close(ld, ldo, 1e-6, 1e-6)


def test_Affine_BatchNorm_golden(golden):
    g = golden("elementwise")
    a = Affine(4)
    z, ld = a(T(g["aff_z_in"]), T(g["aff_params"]))
    close(z, g["aff_z_fwd"], 1e-6, 1e-6); close(ld, g["aff_ld"], 1e-6, 1e-6)
    z, ld = a.inverse_and_log_det(T(g["aff_z_in"]), T(g["aff_params"]))
    close(z, g["aff_z_inv"], 1e-6, 1e-6); close(ld, g["aff_ld_inv"], 1e-6, 1e-6)
    bn = BatchNorm(4, 0.1, 1e-5)
    z, ld = bn(T(g["bn_z_in"]))
    close(bn.get_last_mean(), g["bn_mean"], 1e-5, 1e-5); close(bn.get_last_alpha(), g["bn_alpha"], 1e-4, 1e-6)
    close(z, g["bn_z_fwd"], 1e-4, 2e-4); close(ld, g["bn_ld"], 1e-4, 1e-5)
    z2, ld2 = bn(T(g["bn_z2_in"]), use_last=True)
    close(z2, g["bn_z2_last"], 1e-5, 1e-5); close(ld2, g["bn_ld2"], 1e-4, 1e-5)
    zi, ldi = bn.inverse_and_log_det(T(g["bn_z2_in"]))
    close(zi, g["bn_z_inv"], 1e-5, 1e-5); close(ldi, g["bn_ld_inv"], 1e-4, 1e-5)


def test_BatchNorm():
    """tests/test_bijectors.py:300-346."""
    D = 4
    batch_norm = BatchNorm(D, 0.1, 1e-5)
    assert np.isclose(batch_norm.get_last_mean(), np.zeros(D)).all()
    assert np.isclose(batch_norm.get_last_alpha(), np.ones(D)).all()
    M, N = 20, 50
    z_in = T(np.random.normal(10.0, 1.0, (M, N, D))).float()
    z, log_det = batch_norm(z_in)
    last_mean, last_alpha = batch_norm.get_last_mean(), batch_norm.get_last_alpha()
    z_true = (z_in - last_mean[None, None, :]) / last_alpha[None, None, :]
    assert np.sum((z.numpy() - z_true.numpy()) ** 2) < 1e-2
    assert log_det.dim() == 0
    assert np.isclose(log_det.numpy(), -np.sum(np.log(last_alpha.numpy())))
    z2, log_det = batch_norm(z_in, use_last=True)
    assert np.sum((z2.numpy() - z_true.numpy()) ** 2) < 1e-2
    z_inv, log_det_inv = batch_norm.inverse_and_log_det(z)
    assert np.isclose(log_det.numpy(), log_det_inv.numpy()).all()
    assert np.sum((z_inv.numpy() - z_in.numpy()) ** 2) < 1e-2
    # wide and large: every column, many blocks, deterministic
    for D, rows in ((64, 40000), (300, 3000), (3, 100001)):
        bn = BatchNorm(D)
        zc = (torch.randn(1, rows, D, dtype=torch.float64) * 3.0 + 5.0).float()
        z, ld = bn(zc.cuda())
        zo, ldo, mean, alpha = O.batchnorm_forward(zc.double())
        close(bn.get_last_mean(), mean, 1e-5, 1e-5); close(bn.get_last_alpha(), alpha, 1e-5, 1e-6)
        close(z, zo, 1e-4, 1e-4); close(ld, ldo, 1e-4, 1e-4)
        z_again, _ = bn(zc.cuda())
        assert torch.equal(z, z_again)


def test_ToInterval(golden):
    """tests/test_bijectors.py:203-270 + golden vectors."""
    D = 4
    lb = float("-inf") * np.ones((D,))
    ub = float("inf") * np.ones((D,))
    interval = ToInterval(D, lb, ub)
    M, N = 20, 50
    z_in = T(np.random.normal(0.0, 1.0, (M, N, D)))
    z, log_det = interval(z_in)
    z_inv, log_det_inv = interval.inverse_and_log_det(z)
    assert np.sum((z_in.numpy() - z.numpy()) ** 2) < 1e-10
    assert np.sum((z_in.numpy() - z_inv.numpy()) ** 2) < 1e-10
    assert np.sum((log_det.numpy() - log_det_inv.numpy()) ** 2) < 1e-10

    b = 0.5
    lb = -b * np.array([1., np.inf, 1, np.inf])
    ub = b * np.array([1., 1., np.inf, np.inf])
    interval = ToInterval(D, lb, ub)
    z_in = T(np.random.normal(0.0, 2.0, (M, N, D)))
    z, log_det = interval(z_in)
    assert (z[:, :, 0] > -1).all() and (z[:, :, 0] < 1).all() and (z[:, :, 1] < 1).all() and (z[:, :, 2] > -1).all()
    z_inv, log_det_inv = interval.inverse_and_log_det(z)
    assert np.sum((z_in.numpy() - z_inv.numpy()) ** 2) < 1e-4
    assert np.sum((log_det.numpy() - log_det_inv.numpy()) ** 2) < 1e-4

    interval = ToInterval(D, [-1, -1, -1, -1], np.ones((D,)))
    z, log_det = interval(z_in)
    z_inv, _ = interval.inverse_and_log_det(z)
    assert np.sum((z_in.numpy() - z_inv.numpy()) ** 2) < 1e-10

    g = golden("elementwise")
    ti = ToInterval(6, g["ti_lb"], g["ti_ub"])
    for tag, tol in (("f64", 1e-11), ("f32", 5e-6)):
        z, ld = ti(T(g["ti_%s_z_in" % tag]))
        # fp32 log(1 - tanh(z)^2 + eps) cancels catastrophically for |z| >~ 4: a 1-ulp tanh difference
        # moves the log-det by ~1e-4 there (the reference has the same conditioning)
        close(z, g["ti_%s_z_fwd" % tag], tol, tol); close(ld, g["ti_%s_ld" % tag], tol, tol * 4 if tag == "f64" else 3e-4)
        zi, ldi = ti.inverse_and_log_det(T(g["ti_%s_z_fwd" % tag]))
        zr, lr = g["ti_%s_z_inv" % tag], g["ti_%s_ld_inv" % tag]
        if tag == "f64":
            close(zi, zr, tol, tol); close(ldi, lr, tol, tol * 4)
        else:
            # fp32 atanh / log(exp(x)-1) are ill-conditioned near the interval ends (the reference's own
            # round trip only holds to 1e-4 SSE there): compare rows whose pre-image is moderate
            el = np.abs(zr) < 3.0
            close(zi.numpy()[el], zr[el], 2e-3, 2e-3)
            ok = np.max(np.abs(zr), axis=2) < 3.0
            assert ok.mean() > 0.25
            close(ldi.numpy()[ok], lr[ok], 2e-3, 8e-3)


def test_ToSimplex(golden):
    """tests/test_bijectors.py:349-373 + golden vectors."""
    D = 4
    bij = ToSimplex(D)
    M, N = 20, 50
    z_in = T(np.random.normal(0.0, 1.0, (M, N, D - 1))).float()
    z, log_det = bij(z_in)
    z_np, zin = z.numpy(), z_in.numpy()
    assert np.isclose(np.sum(z_np, 2), 1.0).all()
    expz = np.exp(zin)
    den = np.sum(expz, 2) + 1
    z_true = np.concatenate((expz, np.ones((M, N, 1))), axis=2) / np.expand_dims(den, 2)
    assert np.isclose(z_true, z_np).all()
    g = golden("elementwise")
    z, ld = ToSimplex(int(g["ts_D"]))(T(g["ts_z_in"]))
    close(z, g["ts_z_fwd"], 1e-6, 1e-7); close(ld, g["ts_ld"], 1e-5, 1e-5)
    # wide rows take the warp-per-row kernel
    zc = torch.randn(2, 300, 70)
    z, ld = ToSimplex(71)(zc.cuda())
    zo, ldo = O.tosimplex_forward(zc, 71)
    close(z, zo, 1e-5, 1e-7); close(ld, ldo, 1e-5, 1e-4)


def test_empty_and_ragged_batches():
    b = RealNVP(6, 2, 15)
    P = b.count_num_params()
    for M, N in ((0, 5), (3, 0), (1, 1), (1, 33), (7, 9)):
        z_in = torch.randn(M, N, 6)
        params = torch.randn(M, P) * 0.2
        z, ld = b(z_in, params)
        assert z.shape == (M, N, 6) and ld.shape == (M, N)
        if M * N:
            zo, ldo = O.coupling_forward(z_in, params, 6, 2, 15, True)
            close(z, zo, 1e-5, 1e-5); close(ld, ldo, 1e-5, 1e-5)
    # one shared parameter row against many m (the reference's matmul broadcast)
    z_in = torch.randn(5, 4, 6)
    params = torch.randn(1, P) * 0.2
    z, ld = b(z_in, params)
    zo, ldo = O.coupling_forward(z_in, params, 6, 2, 15, True)
    close(z, zo, 1e-5, 1e-5); close(ld, ldo, 1e-5, 1e-5)


def test_maximum_sizes():
    """Largest conditioner the reference admits: 5 layers of 1000 units."""
    b = RealNVP(9, 5, 1000, transform_upper=False)
    P = b.count_num_params()
    rs = np.random.RandomState(3)
    params = T((rs.standard_normal((2, P)) * 0.03).astype(np.float32))
    z_in = T(rs.standard_normal((2, 3, 9)).astype(np.float32))
    z, ld = b(z_in, params)
    zo, ldo = O.coupling_forward(z_in.double(), params.double(), 9, 5, 1000, False)
    close(z, zo, 1e-4, 1e-4); close(ld, ldo, 1e-4, 1e-4)
    zi, ldi = b.inverse_and_log_det(z, params)
    close(zi, z_in, 1e-4, 1e-4)
