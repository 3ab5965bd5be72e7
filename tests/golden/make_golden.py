"""Generate golden vectors by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Every fixture stores the inputs (or the seed that
regenerates them through torch_nf_b200.synthetic) and the reference outputs.
The reference draws its base noise inside ``NormFlow.forward``
(density_estimator.py:366); it is reproduced here by seeding numpy's legacy
global stream immediately before the call and regenerating the same draw.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import torch_nf.bijectors as rb                     # noqa: E402  (the reference)
import torch_nf.density_estimator as rde            # noqa: E402
import torch_nf.conditional_density_estimator as rcde  # noqa: E402
from torch_nf_b200.synthetic import chain_spec, synthetic_params, synthetic_noise  # noqa: E402

torch.set_num_threads(8)


def save(name, **kw):
    out = {}
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def ref_forward(nf, params, omega_seed, N, freeze_bn=False):
    M = params.shape[0]
    np.random.seed(omega_seed)
    omega = np.random.normal(0.0, 1.0, (M, N, nf.D))
    np.random.seed(omega_seed)
    z, log_q_z = nf.forward(params, N, freeze_bn=freeze_bn)
    return omega, z, log_q_z


def bn_stats(nf):
    means, alphas = [], []
    for b in nf.bijectors:
        if b.name == "BatchNorm":
            means.append(b.get_last_mean().detach().numpy())
            alphas.append(b.get_last_alpha().detach().numpy())
    return np.array(means), np.array(alphas)


# ------------------------------------------------------------------ bijectors
def golden_realnvp():
    rs = np.random.RandomState(0)
    cases = [  # name, D, L, U, upper, M, N, dtype, param std
        ("d4_up_f64", 4, 2, 15, True, 10, 5, np.float64, 0.1),
        ("d4_lo_f64", 4, 2, 15, False, 10, 5, np.float64, 0.1),
        ("d5_lo_f64", 5, 1, 15, False, 20, 5, np.float64, 0.1),
        ("d5_up_f32", 5, 3, 17, True, 7, 3, np.float32, 0.3),
        ("d8_lo_f32", 8, 1, 15, False, 20, 5, np.float32, 0.1),
        ("d8_up_a_f32", 8, 2, 15, True, 1, 300, np.float32, 0.3),   # regime A
        ("d6_up_b_f32", 6, 2, 15, True, 257, 1, np.float32, 0.3),   # regime B
    ]
    out = {}
    for (nm, D, L, U, up, M, N, dt, std) in cases:
        b = rb.RealNVP(D, L, U, transform_upper=up)
        P = b.count_num_params()
        params = torch.tensor((rs.standard_normal((M, P + 3)) * std).astype(dt))  # 3 trailing extras
        z_in = torch.tensor(rs.standard_normal((M, N, D)).astype(dt))
        z, ld = b(z_in, params)
        zi, ldi = b.inverse_and_log_det(z_in, params)
        out.update({nm + "_cfg": np.array([D, L, U, int(up), M, N]), nm + "_params": params,
                    nm + "_z_in": z_in, nm + "_z_fwd": z, nm + "_ld_fwd": ld,
                    nm + "_z_inv": zi, nm + "_ld_inv": ldi})
    # regime A at the headline layer shape; weights by seed only
    D, L, U, M, N = 64, 2, 256, 1, 96
    for up in (True, False):
        b = rb.RealNVP(D, L, U, transform_upper=up)
        params = torch.tensor(synthetic_params([("RealNVP", L, U, up)], D, M, seed=11))
        z_in = torch.tensor(synthetic_noise(M, N, D, seed=12).astype(np.float32))
        z, ld = b(z_in, params)
        zi, ldi = b.inverse_and_log_det(z_in, params)
        nm = "d64_%s_f32" % ("up" if up else "lo")
        out.update({nm + "_cfg": np.array([D, L, U, int(up), M, N]), nm + "_seeds": np.array([11, 12]),
                    nm + "_z_fwd": z, nm + "_ld_fwd": ld, nm + "_z_inv": zi, nm + "_ld_inv": ldi})
    save("realnvp", **out)


def golden_elementwise():
    rs = np.random.RandomState(1)
    out = {}
    # Affine (bijectors.py:277-315)
    D, M, N = 4, 20, 50
    a = rb.Affine(D)
    params = torch.tensor(rs.standard_normal((M, 2 * D)).astype(np.float32))
    z_in = torch.tensor(rs.standard_normal((M, N, D)).astype(np.float32))
    z, ld = a(z_in, params)
    zi, ldi = a.inverse_and_log_det(z_in, params)
    out.update(aff_params=params, aff_z_in=z_in, aff_z_fwd=z, aff_ld=ld, aff_z_inv=zi, aff_ld_inv=ldi)
    # BatchNorm (bijectors.py:389-426)
    bn = rb.BatchNorm(D, 0.1, 1e-5)
    z_in = torch.tensor((10.0 + rs.standard_normal((M, N, D)) * np.array([1.0, 0.1, 3.0, 0.5])).astype(np.float32))
    z, ld = bn(z_in)
    mean, alpha = bn.get_last_mean(), bn.get_last_alpha()
    z2_in = torch.tensor(rs.standard_normal((M, N, D)).astype(np.float32))
    z2, ld2 = bn(z2_in, use_last=True)
    zi, ldi = bn.inverse_and_log_det(z2_in)
    out.update(bn_z_in=z_in, bn_z_fwd=z, bn_ld=ld, bn_mean=mean, bn_alpha=alpha,
               bn_z2_in=z2_in, bn_z2_last=z2, bn_ld2=ld2, bn_z_inv=zi, bn_ld_inv=ldi)
    # ToInterval (bijectors.py:509-557): mixed bounds, both dtypes
    lb = np.array([-0.5, -np.inf, -0.5, -np.inf, 1.0, 0.0])
    ub = np.array([0.5, 0.5, np.inf, np.inf, 4.0, np.inf])
    D = 6
    ti = rb.ToInterval(D, lb, ub)
    for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
        z_in = torch.tensor((rs.standard_normal((M, N, D)) * 2.0).astype(dt))
        z, ld = ti(z_in)
        zi, ldi = ti.inverse_and_log_det(z)
        out.update({"ti_%s_z_in" % tag: z_in, "ti_%s_z_fwd" % tag: z, "ti_%s_ld" % tag: ld,
                    "ti_%s_z_inv" % tag: zi, "ti_%s_ld_inv" % tag: ldi})
    out.update(ti_lb=lb, ti_ub=ub)
    # ToSimplex (bijectors.py:574-591); D attribute is one more than the input width
    D = 4
    ts = rb.ToSimplex(D)
    z_in = torch.tensor(rs.standard_normal((M, N, D - 1)).astype(np.float32))
    z, ld = ts(z_in)
    out.update(ts_z_in=z_in, ts_z_fwd=z, ts_ld=ld, ts_D=np.array(D))
    save("elementwise", **out)


# ------------------------------------------------------------------ NormFlow
def golden_flow(name, D, stages, L, U, M, N, pseed, oseed, support=None, support_tag=None,
                store_params=False, params=None):
    nf = rde.NormFlow(D, True, "coupling", stages, L, U, support)
    spec = chain_spec(nf.bijectors)
    if params is None:
        params = torch.tensor(synthetic_params(spec, D, M, seed=pseed))
    assert params.shape[1] == nf.D_params
    # log_prob before any forward: BatchNorm state is identity (bijectors.py:345-346)
    omega, z, log_q_z = ref_forward(nf, params, oseed, N)
    means, alphas = bn_stats(nf)
    out = dict(cfg=np.array([D, stages, L, U, M, N, pseed, oseed]), z=z, log_q_z=log_q_z,
               bn_mean=means, bn_alpha=alphas)
    if store_params:
        out["params"] = params
        out["omega"] = omega
    if support_tag != "ToSimplex":
        out["log_prob"] = nf.log_prob(z, params)                       # with the stats just stored
        # a second, off-sample evaluation point (also exercises freeze_bn)
        z_b = z + 0.05
        if support_tag == "ToInterval":
            z_b = z
        out["log_prob_b"] = nf.log_prob(z_b, params)
    _, z_f, lq_f = ref_forward(nf, params, oseed + 1, N, freeze_bn=True)
    out["z_frozen"] = z_f
    out["log_q_z_frozen"] = lq_f
    for k in ("z", "log_q_z", "log_prob", "z_frozen"):
        if k in out:
            assert torch.isfinite(out[k]).all(), (name, k)
    save(name, **out)


def golden_flows():
    # C1: 2-D toy (tests/test_density_estimators.py), M=1, N=1024
    golden_flow("flow_c1", 2, 1, 2, 15, 1, 1024, 21, 22, store_params=True)
    # C2a: D=8 shared weights
    golden_flow("flow_c2a", 8, 1, 2, 15, 1, 2048, 23, 24, store_params=True)
    # C2b: per-sample weights produced by the reference's own hyper-network
    torch.manual_seed(0)
    nf = rde.NormFlow(8, True, "coupling", 1, 2, 15)
    cde = rcde.ConditionalDensityEstimator(nf, 8, [100])
    x = torch.tensor(np.random.RandomState(25).standard_normal((48, 8)).astype(np.float32))
    with torch.no_grad():
        params = cde.param_net(x)            # std ~0.28: already far from the identity flow
    weights = {}
    for i, (k, v) in enumerate(cde.param_net.state_dict().items()):
        weights["pn_%s" % k.replace(".", "_")] = v.numpy()
    golden_flow("flow_c2b", 8, 1, 2, 15, 48, 1, 0, 26, store_params=True, params=params)
    save("flow_c2b_net", x=x, params_unscaled=params, **weights)
    # C3 at small N: D=64, 4 stages, U=256 (weights by seed)
    golden_flow("flow_c3", 64, 4, 2, 256, 1, 192, 31, 32)
    # C4-like: ToInterval support, per-sample weights, D=6, U=15
    lb = -2.0 * np.ones(6); ub = 2.0 * np.ones(6)
    golden_flow("flow_c4", 6, 1, 2, 15, 40, 1, 41, 42, support=rb.ToInterval(6, lb, ub),
                support_tag="ToInterval", store_params=True)
    # ToSimplex support (sample direction only), odd D
    golden_flow("flow_simplex", 5, 2, 2, 15, 3, 33, 43, 44, support=rb.ToSimplex(6),
                support_tag="ToSimplex", store_params=True)
    # C5 at small N: D=256, 8 stages (16 coupling layers), U=256
    golden_flow("flow_c5", 256, 8, 2, 256, 1, 64, 51, 52)


def golden_maf():
    """MAF bijector (bijectors.py:597-806) and the 'AR' NormFlow (density_estimator.py:271-274).
    Masks are drawn from numpy's global stream at construction: the seed is stored with them."""
    out = {}
    rs = np.random.RandomState(7)
    for nm, D, L, U, M, N, dt, seed in (("d4_f64", 4, 2, 20, 10, 5, np.float64, 100), ("d20_f64", 20, 2, 20, 3, 4, np.float64, 101),
                                        ("d6_f32", 6, 3, 17, 40, 1, np.float32, 102), ("d5_a_f32", 5, 1, 9, 1, 200, np.float32, 103)):
        np.random.seed(seed)
        b = rb.MAF(D, L, U)
        P = b.count_num_params()
        params = torch.tensor((rs.standard_normal((M, P)) * 0.3).astype(dt))
        z_in = torch.tensor(rs.standard_normal((M, N, D)).astype(dt))
        z, ld = b(z_in, params)
        zi, ldi = b.inverse_and_log_det(z_in, params)
        out.update({nm + "_cfg": np.array([D, L, b.num_units, M, N, seed]), nm + "_params": params, nm + "_z_in": z_in,
                    nm + "_z_fwd": z, nm + "_ld_fwd": ld, nm + "_z_inv": zi, nm + "_ld_inv": ldi})
        for i, Mi in enumerate(b.Ms):
            out["%s_mask%d" % (nm, i)] = Mi
    save("maf", **out)
    # AR flow: MAF -> BatchNorm -> Affine, conditional weights (scripts/lfi_mat.py shape at test size)
    D, L, U, M, N, seed = 6, 2, 20, 32, 4, 104
    np.random.seed(seed)
    nf = rde.NormFlow(D, True, "AR", 1, L, U)
    params = torch.tensor((np.random.RandomState(8).standard_normal((M, nf.D_params)) * 0.3).astype(np.float32))
    omega, z, lq = ref_forward(nf, params, 105, N)
    means, alphas = bn_stats(nf)
    out = dict(cfg=np.array([D, L, U, M, N, seed]), params=params, omega=omega, z=z, log_q_z=lq, bn_mean=means,
               bn_alpha=alphas, log_prob=nf.log_prob(z, params))
    for i, Mi in enumerate(nf.bijectors[0].Ms):
        out["mask%d" % i] = Mi
    save("flow_ar", **out)


if __name__ == "__main__":
    golden_maf()
    sys.exit(0) if "--maf-only" in sys.argv else None
    golden_realnvp()
    golden_elementwise()
    golden_flows()
