"""Host-side API parity (no GPU): constructor validation, clamps, chain
assembly and parameter counts, following the reference's own tests
(tests/test_bijectors.py, test_density_estimators.py,
test_conditional_density_estimators.py, test_error_formatters.py)."""
import numpy as np
import pytest
import torch
from pytest import raises

from oracle import flow_oracle as O
import torch_nf_b200.bijectors as bij
import torch_nf_b200.density_estimator as de
from torch_nf_b200.bijectors import Bijector, RealNVP, MAF, BatchNorm, ToSimplex, Affine, ToInterval
from torch_nf_b200.conditional_density_estimator import ConditionalDensityEstimator
from torch_nf_b200.error_formatters import format_type_err_msg


def test_Bijector_init():
    bijector = Bijector(4)
    assert bijector.D == 4
    with raises(TypeError):
        Bijector("foo")
    with raises(ValueError):
        Bijector(-1)
    z, params = np.zeros((3, 4)), np.zeros((5, 6))
    with raises(NotImplementedError):
        bijector(z, params)
    with raises(NotImplementedError):
        bijector.forward_and_log_det(z, params)
    with raises(NotImplementedError):
        bijector.inverse_and_log_det(z, params)
    assert bijector.count_num_params() == 0


def test_RealNVP_ctor():
    r = RealNVP(4, 2, 15, transform_upper=False)
    assert (r.name, r.D, r.num_layers, r.num_units, r.transform_upper) == ("RealNVP", 4, 2, 15, False)
    r = RealNVP(4, 6, 2000)
    assert r.num_layers == 5 and r.num_units == 1000
    assert RealNVP(4, 3, 10).num_units == 15
    with raises(TypeError):
        RealNVP(4, "foo", 10)
    with raises(ValueError):
        RealNVP(4, -1, 10)
    with raises(TypeError):
        RealNVP(4, 2, "foo")
    with raises(TypeError):
        RealNVP(2, 2, 20, "foo")
    for D, L, U, up in ((4, 2, 15, True), (5, 1, 15, False), (5, 3, 17, True), (64, 2, 256, False), (9, 5, 1000, True)):
        assert RealNVP(D, L, U, up).count_num_params() == O.coupling_num_params(D, L, U, up)
    assert RealNVP(64, 2, 256).count_num_params() == 164928
    assert RealNVP(256, 2, 256).count_num_params() == 263424


def test_MAF_ctor():
    """tests/test_bijectors.py:125-200 (constructor, clamps, masks)."""
    np.random.seed(0)
    maf = MAF(4, 2, 20)
    assert (maf.name, maf.D, maf.num_layers, maf.num_units, maf.fwd_fac) == ("MAF", 4, 2, 20, True)
    assert MAF(4, 6, 2000).num_layers == 5 and MAF(4, 6, 2000).num_units == 1000
    assert MAF(4, 3, 2).num_units == 5
    with raises(TypeError):
        MAF(4, "foo", 10)
    with raises(ValueError):
        MAF(4, -1, 10)
    with raises(TypeError):
        MAF(4, 2, "foo")
    with raises(TypeError):
        MAF(4, 2, 20, "foo")
    assert len(maf.Ms) == 3 and len(maf.ms) == 3
    assert tuple(maf.Ms[0].shape) == (1, 4, 20) and tuple(maf.Ms[1].shape) == (1, 20, 20)
    assert tuple(maf.Ms[2].shape) == (1, 20, 4)
    for m in maf.ms[:-1]:
        assert m.min() >= 1 and m.max() <= 3
    # autoregressive property: output d may only depend on inputs of lower degree
    conn = maf.Ms[0][0] @ maf.Ms[1][0] @ maf.Ms[2][0]
    assert torch.equal(torch.triu(conn, diagonal=1) > 0, conn > 0)
    assert maf.count_num_params() == 2 * (2 * 4 * 20 + 20 * 20)
    assert maf.count_num_params() == O.maf_num_params(4, 2, 20)


def test_MAF_masks_follow_the_reference_stream(golden):
    """Masks are drawn from numpy's global stream in the reference's order (bijectors.py:663-696):
    the same seed reproduces the reference object's masks bit for bit."""
    g = golden("maf")
    for nm in ("d4_f64", "d20_f64", "d6_f32", "d5_a_f32"):
        D, L, U, M, N, seed = [int(v) for v in g[nm + "_cfg"]]
        np.random.seed(seed)
        maf = MAF(D, L, U)
        for i, Mi in enumerate(maf.Ms):
            assert np.array_equal(Mi.numpy(), g["%s_mask%d" % (nm, i)])
        flat = maf._mask_flat(torch.device("cpu"))
        assert flat.numel() == maf.count_num_params()


def test_Affine_BatchNorm_ctor():
    assert Affine(4).count_num_params() == 8 and Affine(4).name == "Affine"
    b = BatchNorm(4, 0.05, 1e-7)
    assert (b.name, b.D, b.momentum, b.eps) == ("BatchNorm", 4, 0.05, 1e-7)
    assert np.isclose(b.get_last_mean(), np.zeros(4)).all()
    assert np.isclose(b.get_last_alpha(), np.ones(4)).all()
    assert BatchNorm(4, 1.01).momentum == 1.0
    with raises(TypeError):
        BatchNorm(4, "foo")
    with raises(ValueError):
        BatchNorm(4, -1.0)
    with raises(TypeError):
        BatchNorm(4, 0.5, "foo")
    with raises(ValueError):
        BatchNorm(4, 0.5, -1.0)
    assert BatchNorm(4, 0.1, 1e-5).count_num_params() == 0


def test_ToInterval_ToSimplex_ctor():
    D = 4
    with raises(ValueError):
        ToInterval(D, -np.ones((D,)), np.ones((D + 1,)))
    ub = np.ones((D,)); ub[3] = -2
    with raises(ValueError):
        ToInterval(D, -np.ones((D,)), ub)
    with raises(TypeError):
        ToInterval(D, "[-1,-1,-1,-1]", np.ones((D,)))
    with raises(TypeError):
        ToInterval(D, -np.ones((D,)), "[1,1,1,1]")
    ti = ToInterval(D, [-1, -np.inf, 0, -np.inf], [1, 2, np.inf, np.inf])
    c = O.tointerval_consts([-1, -np.inf, 0, -np.inf], [1, 2, np.inf, np.inf])
    assert torch.equal(ti.tanh_flg, c["tanh_flg"]) and torch.equal(ti.softplus_flg, c["sp_flg"])
    assert torch.equal(ti.tanh_m, c["tanh_m"]) and torch.equal(ti.tanh_c, c["tanh_c"])
    assert torch.equal(ti.softplus_m, c["sp_m"]) and torch.equal(ti.softplus_c, c["sp_c"])
    assert ti.name == "ToInterval" and ti.count_num_params() == 0
    ts = ToSimplex(4)
    assert ts.name == "ToSimplex" and ts.D == 4 and ts.count_num_params() == 0


def test_DensityEstimator():
    d = de.DensityEstimator(4, False)
    assert d.D == 4 and not d.conditioner
    with raises(TypeError):
        de.DensityEstimator("foo", False)
    with raises(ValueError):
        de.DensityEstimator(1, False)
    with raises(TypeError):
        de.DensityEstimator(4, "foo")
    with raises(NotImplementedError):
        d.forward(None)
    with raises(NotImplementedError):
        d.log_prob(None)
    with raises(NotImplementedError):
        d.count_num_params()
    with raises(NotImplementedError):
        d._param_init()


def test_NormFlow_ctor():
    """tests/test_density_estimators.py:147-224."""
    D = 4
    nf = de.NormFlow(D, False, "coupling", 1, 2, 30, None)
    assert (nf.arch_type, nf.num_stages, nf.num_layers, nf.num_units, nf.support_layer) == ("coupling", 1, 2, 30, None)
    nf = de.NormFlow(D, False, "coupling", 1, 2, 10, bij.ToSimplex(D))
    assert nf.num_units == 15 and issubclass(type(nf.support_layer), bij.Bijector)
    bad = [(TypeError, ("foo", False, "coupling", 1, 2, 20, None)), (ValueError, (-1, False, "coupling", 1, 2, 20, None)),
           (TypeError, (4, False, 1, 1, 2, 20, None)), (ValueError, (4, False, "foo", 1, 2, 20, None)),
           (TypeError, (4, 1, "coupling", 1, 2, 20, None)), (TypeError, (4, False, "coupling", "foo", 2, 20, None)),
           (ValueError, (4, False, "coupling", -1, 2, 20, None)), (TypeError, (4, False, "coupling", 1, "foo", 20, None)),
           (ValueError, (4, False, "coupling", 1, -1, 20, None)), (TypeError, (4, False, "coupling", 1, 2, "foo", None)),
           (ValueError, (4, False, "coupling", 1, 2, -1, None)), (TypeError, (4, False, "coupling", 1, 2, 20, "foo"))]
    for exc, args in bad:
        with raises(exc):
            de.NormFlow(*args)
    nf = de.NormFlow(D, True, "coupling", 2, 2, 20, bij.ToSimplex(D))
    kinds = [bij.RealNVP, bij.BatchNorm, bij.RealNVP, bij.BatchNorm, bij.Affine] * 2 + [bij.ToSimplex]
    assert [type(b) for b in nf.bijectors] == kinds
    assert [b.transform_upper for b in nf.bijectors if b.name == "RealNVP"] == [True, False, True, False]
    assert issubclass(type(de.NormFlow(D, False, "AR", num_layers=2, num_units=20).bijectors[0]), bij.MAF)
    assert issubclass(type(de.NormFlow(D, False, "affine").bijectors[0]), bij.Affine)


def test_NormFlow_param_counts_and_init():
    # C1 1148, C2 1532, C3 1319936, C5 4218880
    for args, n in (((2, True, "coupling", 1, 2, 15), 1148), ((8, True, "coupling", 1, 2, 15), 1532),
                    ((64, True, "coupling", 4, 2, 256), 1319936), ((256, True, "coupling", 8, 2, 256), 4218880)):
        nf = de.NormFlow(*args)
        assert nf.D_params == n
        chain = O.build_chain(args[0], "coupling", args[3], args[4], args[5])
        assert O.chain_num_params(chain, args[0]) == n
        assert not hasattr(nf, "params")
    torch.manual_seed(0)
    nf = de.NormFlow(4, False, "coupling", 1, 2, 20)
    assert tuple(nf.params.shape) == (1, nf.D_params) and nf.params.requires_grad and nf.params.is_leaf
    assert abs(float(nf.params.std()) - np.sqrt(2.0 / (nf.D_params + 1))) < 0.1 * np.sqrt(2.0 / (nf.D_params + 1))
    # parameter slicing is in chain order
    offs = [(b.name, i, n) for (b, i, n) in nf._slices()]
    assert offs[0] == ("RealNVP", 0, nf.bijectors[0].count_num_params()) and offs[1][2] == 0
    assert offs[-1] == ("Affine", nf.D_params - 8, 8)


def test_ConditionalDensityEstimator_ctor():
    """tests/test_conditional_density_estimators.py:32-87."""
    nf = de.NormFlow(4, True, "coupling", 1, 2, 20)
    cde = ConditionalDensityEstimator(nf, 3, [10, 12], dropout=True)
    assert cde.D_x == 3 and cde.D_params == nf.D_params and cde.hidden_layers == [10, 12]
    assert list(dict(cde.param_net.named_children()).keys()) == [
        "linear1", "tanh1", "dropout1", "linear2", "relu2", "dropout2", "linear3"]
    assert isinstance(cde.param_net.relu2, torch.nn.Tanh)
    assert cde.param_net.linear3.out_features == nf.D_params
    with raises(TypeError):
        ConditionalDensityEstimator("foo", 3, [10])
    with raises(TypeError):
        ConditionalDensityEstimator(nf, "foo", [10])
    with raises(ValueError):
        ConditionalDensityEstimator(nf, 0, [10])
    with raises(TypeError):
        ConditionalDensityEstimator(nf, 3, "foo")
    with raises(TypeError):
        ConditionalDensityEstimator(nf, 3, [10, "foo"])
    with raises(ValueError):
        ConditionalDensityEstimator(nf, 3, [10, 0])
    nf.D_params = 4.
    with raises(TypeError):
        ConditionalDensityEstimator(nf, 3, [10])
    nf.D_params = 0
    with raises(ValueError):
        ConditionalDensityEstimator(nf, 3, [10])

    class Sub(de.NormFlow):
        pass
    with raises(TypeError):                       # exact type only, like the reference
        ConditionalDensityEstimator(Sub(4, True, "coupling", 1, 2, 20), 3, [10])


def test_error_formatters():
    x = 20
    assert format_type_err_msg(BatchNorm(2), "foo", x, float) == "BatchNorm argument foo must be float not int."
    with raises(ValueError):
        format_type_err_msg(BatchNorm(2), "foo", 1.0, float)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    """The product path must fail loudly, never compute on the CPU."""
    nf = de.NormFlow(4, False, "coupling", 1, 2, 20)
    with raises(RuntimeError, match="no CPU fallback"):
        nf(10)
    with raises(RuntimeError, match="no CPU fallback"):
        nf.log_prob(torch.zeros(1, 3, 4))
    for b, args in ((RealNVP(4, 2, 15), (torch.zeros(1, 2, 4), torch.zeros(1, 600))), (Affine(4), (torch.zeros(1, 2, 4), torch.zeros(1, 8))),
                    (BatchNorm(4), (torch.zeros(1, 2, 4),)), (ToSimplex(4), (torch.zeros(1, 2, 3),))):
        with raises(RuntimeError, match="no CPU fallback"):
            b(*args)


def test_training_backward_switch():
    """config.set_training_backward: validation and the mode table (no GPU needed)."""
    from torch_nf_b200 import config
    old_p, old_b = config.conditioner_precision(), config.training_backward()
    try:
        with pytest.raises(ValueError):
            config.set_training_backward("fp16")
        table = {("fp32", "auto"): False, ("bf16", "auto"): True, ("fp32_cc", "auto"): False,
                 ("fp32", "bf16"): True, ("bf16", "bf16"): True, ("fp32_cc", "bf16"): False,
                 ("fp32", "exact"): False, ("bf16", "exact"): False}
        for (prec, bwd), want in table.items():
            config.set_conditioner_precision(prec)
            config.set_training_backward(bwd)
            assert config.tc_backward_enabled() is want, (prec, bwd)
    finally:
        config.set_conditioner_precision(old_p)
        config.set_training_backward(old_b)
