"""Pin the CPU oracle (oracle/flow_oracle.py) to golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

from oracle import flow_oracle as O
from torch_nf_b200.synthetic import synthetic_params, synthetic_noise

T = torch.tensor


def close(a, b, rtol, atol):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


REALNVP_CASES = ["d4_up_f64", "d4_lo_f64", "d5_lo_f64", "d5_up_f32", "d8_lo_f32", "d8_up_a_f32", "d6_up_b_f32"]


def test_realnvp_small(golden):
    g = golden("realnvp")
    for nm in REALNVP_CASES:
        D, L, U, up, M, N = [int(v) for v in g[nm + "_cfg"]]
        params, z_in = T(g[nm + "_params"]), T(g[nm + "_z_in"])
        tol = 1e-12 if params.dtype == torch.float64 else 2e-6
        z, ld = O.coupling_forward(z_in, params, D, L, U, bool(up))
        close(z, g[nm + "_z_fwd"], tol, tol); close(ld, g[nm + "_ld_fwd"], tol, tol)
        z, ld = O.coupling_inverse(z_in, params, D, L, U, bool(up))
        close(z, g[nm + "_z_inv"], tol, tol); close(ld, g[nm + "_ld_inv"], tol, tol)
        assert O.coupling_num_params(D, L, U, bool(up)) == params.shape[1] - 3


def test_realnvp_headline_shape(golden):
    g = golden("realnvp")
    for nm, up in (("d64_up_f32", True), ("d64_lo_f32", False)):
        D, L, U, _, M, N = [int(v) for v in g[nm + "_cfg"]]
        ps, zs = [int(v) for v in g[nm + "_seeds"]]
        params = T(synthetic_params([("RealNVP", L, U, up)], D, M, seed=ps))
        z_in = T(synthetic_noise(M, N, D, seed=zs).astype(np.float32))
        z, ld = O.coupling_forward(z_in, params, D, L, U, up)
        close(z, g[nm + "_z_fwd"], 1e-5, 1e-5); close(ld, g[nm + "_ld_fwd"], 1e-5, 1e-5)
        z, ld = O.coupling_inverse(z_in, params, D, L, U, up)
        close(z, g[nm + "_z_inv"], 1e-5, 1e-5); close(ld, g[nm + "_ld_inv"], 1e-5, 1e-5)


def test_affine_batchnorm(golden):
    g = golden("elementwise")
    z, ld = O.affine_forward(T(g["aff_z_in"]), T(g["aff_params"]), 4)
    close(z, g["aff_z_fwd"], 1e-6, 1e-6); close(ld, g["aff_ld"], 1e-6, 1e-6)
    assert tuple(ld.shape) == (20, 1)
    z, ld = O.affine_inverse(T(g["aff_z_in"]), T(g["aff_params"]), 4)
    close(z, g["aff_z_inv"], 1e-6, 1e-6); close(ld, g["aff_ld_inv"], 1e-6, 1e-6)

    z, ld, mean, alpha = O.batchnorm_forward(T(g["bn_z_in"]), 1e-5)
    close(mean, g["bn_mean"], 1e-5, 1e-5); close(alpha, g["bn_alpha"], 1e-4, 1e-6)
    close(z, g["bn_z_fwd"], 1e-4, 2e-4); close(ld, g["bn_ld"], 1e-4, 1e-5)
    assert ld.dim() == 0
    z2, ld2, _, _ = O.batchnorm_forward(T(g["bn_z2_in"]), 1e-5, True, T(g["bn_mean"]), T(g["bn_alpha"]))
    close(z2, g["bn_z2_last"], 1e-6, 1e-6); close(ld2, g["bn_ld2"], 1e-6, 1e-6)
    zi, ldi = O.batchnorm_inverse(T(g["bn_z2_in"]), T(g["bn_mean"]), T(g["bn_alpha"]))
    close(zi, g["bn_z_inv"], 1e-6, 1e-6); close(ldi, g["bn_ld_inv"], 1e-6, 1e-6)


def test_support_layers(golden):
    g = golden("elementwise")
    lb, ub = g["ti_lb"], g["ti_ub"]
    for tag, tol in (("f64", 1e-12), ("f32", 2e-6)):
        z, ld = O.tointerval_forward(T(g["ti_%s_z_in" % tag]), lb, ub)
        close(z, g["ti_%s_z_fwd" % tag], tol, tol); close(ld, g["ti_%s_ld" % tag], tol, tol)
        zi, ldi = O.tointerval_inverse(T(g["ti_%s_z_fwd" % tag]), lb, ub)
        close(zi, g["ti_%s_z_inv" % tag], tol, tol); close(ldi, g["ti_%s_ld_inv" % tag], tol, tol)
    z, ld = O.tosimplex_forward(T(g["ts_z_in"]), int(g["ts_D"]))
    close(z, g["ts_z_fwd"], 1e-6, 1e-7); close(ld, g["ts_ld"], 1e-6, 1e-6)


def _flow_case(golden, name, support=None, has_logprob=True, tol=2e-5):
    g = golden(name)
    D, stages, L, U, M, N, pseed, oseed = [int(v) for v in g["cfg"]]
    chain = O.build_chain(D, "coupling", stages, L, U, support)
    if "params" in g.files:
        params, omega = T(g["params"]), g["omega"]
    else:
        from torch_nf_b200.synthetic import chain_spec  # noqa: F401
        spec = [(b["kind"], b.get("L", 0), b.get("U", 0), b.get("upper", False)) for b in chain]
        params = T(synthetic_params(spec, D, M, seed=pseed))
        np.random.seed(oseed)
        omega = np.random.normal(0.0, 1.0, (M, N, D))
    assert O.chain_num_params(chain, D) == params.shape[1]
    z, lq, st = O.normflow_forward(chain, D, params, omega)
    assert z.dtype == torch.float32 and lq.dtype == torch.float64
    close(z, g["z"], tol, tol)
    close(lq, g["log_q_z"], tol, tol * 10)
    means = np.array([m.numpy() for (m, a) in [s for s in st if s is not None]])
    alphas = np.array([a.numpy() for (m, a) in [s for s in st if s is not None]])
    close(means, g["bn_mean"], tol, tol); close(alphas, g["bn_alpha"], tol, tol)
    gst = []
    it = iter(range(len(g["bn_mean"])))
    for b in chain:
        if b["kind"] == "BatchNorm":
            i = next(it)
            gst.append((T(g["bn_mean"][i]), T(g["bn_alpha"][i])))
        else:
            gst.append(None)
    if has_logprob:
        lp = O.normflow_log_prob(chain, D, T(g["z"]), params, gst)
        assert lp.dtype == torch.float32
        close(lp, g["log_prob"], tol, tol * 10)
    np.random.seed(oseed + 1)
    omega2 = np.random.normal(0.0, 1.0, (M, N, D))
    z_f, lq_f, _ = O.normflow_forward(chain, D, params, omega2, freeze_bn=True, bn_state=gst)
    close(z_f, g["z_frozen"], tol, tol); close(lq_f, g["log_q_z_frozen"], tol, tol * 10)


def test_flow_c1(golden):
    _flow_case(golden, "flow_c1")


def test_flow_c2(golden):
    _flow_case(golden, "flow_c2a")
    _flow_case(golden, "flow_c2b")


def test_flow_c3_c5(golden):
    _flow_case(golden, "flow_c3", tol=5e-5)
    _flow_case(golden, "flow_c5", tol=1e-4)


def test_flow_support(golden):
    lb = -2.0 * np.ones(6); ub = 2.0 * np.ones(6)
    _flow_case(golden, "flow_c4", support=("ToInterval", lb, ub), tol=5e-5)
    _flow_case(golden, "flow_simplex", support=("ToSimplex", 6), has_logprob=False)


def test_param_net(golden):
    g = golden("flow_c2b_net")
    ws = [(T(g["pn_linear1_weight"]), T(g["pn_linear1_bias"])), (T(g["pn_linear2_weight"]), T(g["pn_linear2_bias"]))]
    close(O.param_net(T(g["x"]), ws), g["params_unscaled"], 1e-5, 1e-6)


MAF_CASES = ["d4_f64", "d20_f64", "d6_f32", "d5_a_f32"]


def test_maf(golden):
    g = golden("maf")
    for nm in MAF_CASES:
        D, L, U, M, N, seed = [int(v) for v in g[nm + "_cfg"]]
        masks = [T(g["%s_mask%d" % (nm, i)]) for i in range(L + 1)]
        params, z_in = T(g[nm + "_params"]), T(g[nm + "_z_in"])
        tol = 1e-11 if params.dtype == torch.float64 else 5e-6
        assert O.maf_num_params(D, L, U) == params.shape[1]
        z, ld = O.maf_forward(z_in, params, masks, D, L, U)
        close(z, g[nm + "_z_fwd"], tol, tol); close(ld, g[nm + "_ld_fwd"], tol, tol)
        z, ld = O.maf_inverse(z_in, params, masks, D, L, U)
        close(z, g[nm + "_z_inv"], tol, tol); close(ld, g[nm + "_ld_inv"], tol, tol)
