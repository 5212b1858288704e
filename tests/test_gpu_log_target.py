"""GPU parity: chain-batched log_target / gradient kernel vs the reference's golden vectors and the oracle.
Tolerances are BASELINE.json's: 1e-10 relative at fp64, 1e-5 at fp32.  The fp32 kernels are measured against fp64 truth:
the unmodified reference evaluated in fp64 at the fp32 fixtures (tests/golden/model_goldens_f32ref.npz) -- torch's own fp32
evaluation of the same fixtures (model_goldens.npz) is itself up to 2.7e-6 away from it."""
import numpy as np
import pytest
import torch

import oracle
from gpu_helpers import dataset, make_model, npy, T_DTYPES
from helpers import ARCHS, NP_DTYPES, PRIOR_SCALES, RTOL, data_of, load, rel_err, spec_of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("arch", list(ARCHS))
@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("pst", list(PRIOR_SCALES))
@pytest.mark.parametrize("tt,temp", [("", None), ("_T07", 0.7)])
def test_batched_eval_vs_reference_goldens(arch, tag, pst, tt, temp):
    mg = load("model_goldens")
    key = f"{arch}_{tag}_{pst}{tt}"
    m = make_model(arch, tag, PRIOR_SCALES[pst], temp)
    ds = dataset(arch, tag)
    theta = torch.from_numpy(mg[key + "_theta"]).to(T_DTYPES[tag])
    tol = RTOL[tag]
    if tag == "f64":
        want_lt, want_g = mg[key + "_lt"], mg[key + "_grad"]
    else:
        ref64 = load("model_goldens_f32ref")
        want_lt, want_g = ref64[key + "_lt64"], ref64[key + "_grad64"]
    for lanes in (0, 1, 4, 8, 16, 32):
        lt, g = m.upto_grad_log_target_batch(theta, ds.x, ds.y, lanes=lanes)
        lt, g = npy(lt), npy(g)
        assert np.allclose(lt, want_lt, rtol=tol, atol=0), (lanes, lt, want_lt)
        for c in range(theta.shape[0]):
            assert rel_err(g[c], want_g[c]) < tol, (lanes, c)
        lt_only = npy(m.log_target_batch(theta, ds.x, ds.y, lanes=lanes))
        assert np.array_equal(lt_only, lt)


@pytest.mark.parametrize("arch", list(ARCHS))
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_single_chain_api_matches_reference(arch, tag):
    """The reference's per-call surface: log_target, grad_log_target, upto_grad_log_target, log_lik, log_prior, forward."""
    mg = load("model_goldens")
    key = f"{arch}_{tag}_p100"
    m = make_model(arch, tag, 100.0)
    ds = dataset(arch, tag)
    tol = RTOL[tag]
    sfx = ""
    if tag == "f32":                      # fp64 truth at the fp32 fixtures
        mg_v, sfx = load("model_goldens_f32ref"), "64"
    else:
        mg_v = mg
    theta = torch.from_numpy(mg[key + "_theta"][0]).to(T_DTYPES[tag])
    lt = m.log_target(theta.clone().detach(), ds.x, ds.y)
    assert lt.dim() == 0 and abs(lt.item() - mg_v[key + "_lt" + sfx][0]) <= tol * abs(mg_v[key + "_lt" + sfx][0])
    g = m.grad_log_target(lt)
    assert g.shape == (m.num_params(),) and rel_err(npy(g), mg_v[key + "_grad" + sfx][0]) < tol
    lt2, g2 = m.upto_grad_log_target(theta.clone().detach(), ds.x, ds.y)
    assert lt2.item() == lt.item() and torch.equal(g2, g)
    for name, got in (("_ll", m.log_lik(ds.x, ds.y).item()), ("_lp", m.log_prior().item())):
        want = mg_v[key + name + sfx][0]
        bar = tol
        if tag == "f32":
            # the naive BCE on fp32 probabilities (stats/loss.py:2) is ill-conditioned at saturated units: for the 2-2-1 fixture
            # torch's own fp32 log_lik is 1.1e-5 from the fp64 truth.  In the reference dtype the bar is 1e-5 (north_star);
            # against the fp64 truth: 1e-5, or twice the reference's own fp32 error where that is larger
            assert abs(got - mg[key + name][0]) <= tol * abs(mg[key + name][0]), name
            bar = max(tol, 2 * abs(mg[key + name][0] - want) / abs(want))
        assert abs(got - want) <= bar * abs(want), name
    assert torch.equal(m.get_params(), theta.to(m.device))
    out = npy(m(ds.x))
    x, y = data_of(arch, NP_DTYPES[tag], mg)
    ref = oracle.forward(spec_of(arch), npy(theta)[None], x)[-1][0]
    assert out.shape == ref.shape and rel_err(out, ref) < (1e-12 if tag == "f64" else 1e-5)


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_saturation_gives_nan_like_the_reference(tag):
    mg = load("model_goldens")
    m = make_model("221", tag, 1.0)
    ds = dataset("221", tag)
    th = torch.from_numpy(mg[f"sat_{tag}_theta"]).to(T_DTYPES[tag])
    lt, g = m.upto_grad_log_target(th, ds.x, ds.y)
    assert torch.isnan(lt) and torch.isnan(g).all()
    assert np.isnan(mg[f"sat_{tag}_lt"])


@pytest.mark.parametrize("arch,tag,n_rows", [("433", "f64", 150), ("2321", "f64", 200), ("2321", "f32", 203),
                                              ("4323", "f32", 77), ("221", "f64", 1), ("433", "f64", 3)])
def test_random_data_vs_oracle_ragged_rows(arch, tag, n_rows):
    """Synthetic data of the BASELINE shapes (iris-shaped N=150, noisy-XOR-shaped N=200) and ragged row counts that do
    not divide the lane count or the 16-byte bulk-copy granule."""
    rng = np.random.default_rng(n_rows)
    dt = NP_DTYPES[tag]
    spec = spec_of(arch)
    d0, k = spec.dims[0], spec.dims[-1]
    x = rng.normal(size=(n_rows, d0)).astype(dt)
    if k == 1:
        y = rng.integers(0, 2, size=(n_rows, 1)).astype(dt)
    else:
        y = np.eye(k, dtype=dt)[rng.integers(0, k, size=n_rows)]
    C = 133
    theta = (rng.normal(size=(C, spec.num_params)) * 0.8).astype(dt)
    loc = rng.normal(size=spec.num_params).astype(dt) * 0.1
    scale = (0.5 + rng.uniform(size=spec.num_params)).astype(dt)
    m = make_model(arch, tag)
    m.prior = torch.distributions.Normal(torch.from_numpy(loc), torch.from_numpy(scale))
    lt_ref, g_ref = oracle.log_target_grad(spec, theta.astype(np.float64), x.astype(np.float64), y,
                                           loc.astype(np.float64), scale.astype(np.float64))
    tol = RTOL[tag]                       # 1e-10 (fp64) / 1e-5 (fp32) against the fp64 oracle, as BASELINE.json states
    for lanes in (1, 4, 8, 16, 32):
        lt, g = m.upto_grad_log_target_batch(torch.from_numpy(theta), torch.from_numpy(x), torch.from_numpy(y), lanes=lanes)
        assert np.allclose(npy(lt), lt_ref, rtol=tol, atol=0), lanes
        for c in range(C):
            assert rel_err(npy(g)[c], g_ref[c]) < tol, (lanes, c)


def test_unaligned_data_pointer_takes_the_plain_load_path():
    m = make_model("2321", "f64")
    ds = dataset("2321", "f64")
    buf = torch.zeros(9, dtype=torch.float64, device="cuda")
    x_un = buf[1:9].view(4, 2)          # 8-byte aligned only
    x_un.copy_(ds.x)
    theta = torch.randn(5, 20, dtype=torch.float64)
    a, ga = m.upto_grad_log_target_batch(theta, ds.x, ds.y)
    b, gb = m.upto_grad_log_target_batch(theta, x_un, ds.y)
    assert torch.equal(a, b) and torch.equal(ga, gb)


def test_errors():
    m = make_model("221", "f64")
    ds = dataset("221", "f64")
    with pytest.raises(ValueError):
        m.log_target(torch.zeros(9, dtype=torch.float64), ds.x[:, :1], ds.y)
    with pytest.raises(ValueError):
        m.set_params(torch.zeros(4, dtype=torch.float64))
    from eeyore_b200.constants import loss_functions
    from eeyore_b200.models.mlp import MLP, Hyperparameters
    bad = MLP(loss=loss_functions["binary_classification"], hparams=Hyperparameters([3, 5, 1], activations=[torch.sigmoid, None]))
    with pytest.raises(ValueError, match="sigmoid output"):
        bad.log_target(torch.zeros(bad.num_params(), dtype=torch.float64), torch.zeros(2, 3, dtype=torch.float64),
                       torch.zeros(2, 1, dtype=torch.float64))
    with pytest.raises(ValueError):
        MLP(loss=loss_functions["binary_classification"], hparams=Hyperparameters([3, 5, 1], activations=[torch.tanh, torch.sigmoid])
            ).log_target(torch.zeros(26, dtype=torch.float64), torch.zeros(2, 3, dtype=torch.float64), torch.zeros(2, 1, dtype=torch.float64))


def test_full_size_chain_batch_properties():
    """BASELINE config 4 size (524,288 chains, 2-3-2-1, XOR): size-independent properties -- log_target = log_lik +
    log_prior, invariance to the lane mapping, and a random subset against the oracle."""
    m = make_model("2321", "f64", 3 ** 0.5)
    ds = dataset("2321", "f64")
    C = 524288
    g = torch.Generator(device="cuda").manual_seed(0)
    theta = torch.randn(C, 20, dtype=torch.float64, device="cuda", generator=g) * 1.7
    xd, yd = m._to_dev(ds.x), m._to_dev(ds.y)
    lt, gr, ll, lp = m._eval(theta, xd, yd, want_grad=True, parts=True)
    assert torch.equal(lt, ll + lp)
    lt4, gr4 = m._eval(theta, xd, yd, want_grad=True, lanes=4)
    assert torch.allclose(lt, lt4, rtol=1e-13, atol=0) and torch.allclose(gr, gr4, rtol=1e-11, atol=1e-13)
    idx = torch.randint(0, C, (257,), generator=torch.Generator().manual_seed(1))
    lt_ref, g_ref = oracle.log_target_grad(spec_of("2321"), npy(theta[idx.cuda()]), npy(ds.x), npy(ds.y),
                                           np.zeros(20), np.full(20, 3 ** 0.5))
    assert np.allclose(npy(lt[idx.cuda()]), lt_ref, rtol=1e-10, atol=0)
    assert rel_err(npy(gr[idx.cuda()]), g_ref) < 1e-10


def test_fp64_fast_sigmoid_accuracy():
    """The fp64 kernels use a table-based exp (argument reduction by one FMA in base 2: its error grows like |a| 2^-53)
    and a refined MUFU reciprocal; outputs stay within ~2 ulp + |a| / 2 ulp of 1/(1+exp(-a)) -- five orders of magnitude
    under the 1e-10 parity bar at |a| = 700 -- and saturate exactly like the reference (p == 1.0 for a >= 36.8)."""
    m = make_model("221", "f64")
    b = np.concatenate([np.linspace(-700, 700, 4001), np.linspace(-40, 40, 8001), np.random.default_rng(0).normal(size=4000) * 3])
    theta = np.zeros((b.size, 9))
    theta[:, 8] = b                                     # all weights 0: output = sigmoid(b2)
    x = torch.zeros(1, 2, dtype=torch.float64)
    out = npy(m.forward_batch(torch.from_numpy(theta), x))[:, 0, 0]
    with np.errstate(over="ignore"):
        ref = 1.0 / (1.0 + np.exp(-b))
    ok = ref > 1e-290
    rel = np.abs(out[ok] - ref[ok]) / ref[ok]
    assert np.all(rel < 6e-16 + 1.2e-16 * np.abs(b[ok])), rel.max()
    assert rel[np.abs(b[ok]) <= 3].max() < 8e-16
    assert np.all(out[b >= 37.0] == 1.0) and np.all(out[b <= -37] < 1e-15) and np.all(out > 0)
    # hidden-layer use: sigmoid(w * x + b) through a full evaluation stays within the 1e-10 parity bar trivially; here the
    # tighter check is on the gradient of a saturating unit
    ds = dataset("221", "f64")
    th = torch.from_numpy(np.random.default_rng(1).normal(size=(64, 9)) * 6)
    lt, g = m.upto_grad_log_target_batch(th, ds.x, ds.y)
    lt_ref, g_ref = oracle.log_target_grad(spec_of("221"), npy(th), npy(ds.x), npy(ds.y), np.zeros(9), np.ones(9))
    fin = np.isfinite(lt_ref)
    assert np.allclose(npy(lt)[fin], lt_ref[fin], rtol=1e-12, atol=0)
    assert np.array_equal(np.isnan(npy(lt)), np.isnan(lt_ref))


GEN_CASES = [
    ([3, 5, 1], [True, True], [torch.sigmoid, torch.sigmoid], "binary_classification"),
    ([4, 3, 3], [True, False], [torch.sigmoid, None], "multiclass_classification"),
    ([4, 6, 5, 4, 3], [True, False, True, True], [torch.sigmoid, None, torch.sigmoid, None], "multiclass_classification"),
    ([2, 4, 1], [False, False], [None, torch.sigmoid], "binary_classification"),
    ([5, 7, 3, 2, 4, 1], [True] * 5, [torch.sigmoid] * 5, "binary_classification"),
    ([10, 16, 16, 1], [True] * 3, [torch.sigmoid] * 3, "binary_classification"),
]


@pytest.mark.parametrize("dims,bias,acts,loss", GEN_CASES)
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_runtime_shape_networks_vs_oracle(dims, bias, acts, loss, tag):
    """mlp.Hyperparameters in full generality (dims, per-layer bias flags, sigmoid / None activations; mlp.py:9-19,37-50)
    through the runtime-shape kernels, incl. samplers."""
    from eeyore_b200.constants import loss_functions
    from eeyore_b200.models.mlp import MLP, Hyperparameters
    from eeyore_b200.samplers import HMC, MALA, MetropolisHastings
    from eeyore_b200.datasets import XYDataset
    from gpu_helpers import loader
    from oracle.mlp import MLPSpec
    dt, tdt = NP_DTYPES[tag], T_DTYPES[tag]
    spec = MLPSpec(dims, loss=loss, bias=bias, activations=["sigmoid" if a is not None else None for a in acts])
    rng = np.random.default_rng(sum(dims))
    n = 37
    x = rng.normal(size=(n, dims[0])).astype(dt)
    k = dims[-1]
    y = (rng.integers(0, 2, size=(n, 1)).astype(dt) if k == 1 else np.eye(k, dtype=dt)[rng.integers(0, k, size=n)])
    m = MLP(loss=loss_functions[loss], hparams=Hyperparameters(dims, bias, acts), dtype=tdt)
    P = m.num_params()
    assert P == spec.num_params
    loc = (rng.normal(size=P) * 0.1).astype(dt); scale = (0.5 + rng.uniform(size=P)).astype(dt)
    m.prior = torch.distributions.Normal(torch.from_numpy(loc), torch.from_numpy(scale))
    C = 70
    theta = (rng.normal(size=(C, P)) * 0.5).astype(dt)
    lt, g = m.upto_grad_log_target_batch(torch.from_numpy(theta), torch.from_numpy(x), torch.from_numpy(y))
    lt_ref, g_ref = oracle.log_target_grad(spec, theta.astype(np.float64), x.astype(np.float64), y, loc.astype(np.float64),
                                           scale.astype(np.float64))
    tol = RTOL[tag]
    assert np.allclose(npy(lt), lt_ref, rtol=tol, atol=0)
    for c in range(C):
        assert rel_err(npy(g)[c], g_ref[c]) < tol
    out = npy(m.forward_batch(torch.from_numpy(theta[:3]), torch.from_numpy(x)))
    ref_out = oracle.forward(spec, theta[:3].astype(np.float64), x.astype(np.float64))[-1]
    assert rel_err(out, ref_out) < (1e-12 if tag == "f64" else 1e-5)
    if tag == "f32":
        return
    T, nb = 8, 2
    z, u = rng.normal(size=(T, C, P)), rng.uniform(size=(T, C))
    ds = XYDataset(torch.from_numpy(x), torch.from_numpy(y))
    for cls, kw, ref_fn in ((MetropolisHastings, dict(scale=0.05), lambda: oracle.mh_run(spec, x, y, loc, scale, theta, z, u, n_burnin=nb, prop_scale=0.05)),
                            (MALA, dict(step=0.01), lambda: oracle.mala_run(spec, x, y, loc, scale, theta, z, u, 0.01, n_burnin=nb)),
                            (HMC, dict(step=0.03, num_steps=3), lambda: oracle.hmc_run(spec, x, y, loc, scale, theta, z, u, 0.03, 3, n_burnin=nb))):
        s = cls(m, theta0=torch.from_numpy(theta), dataloader=loader(ds), **kw)
        s.set_noise_tape(torch.from_numpy(z), torch.from_numpy(u))
        s.run(num_epochs=T, num_burnin_epochs=nb)
        got, ref = s.get_chain(), ref_fn()
        assert np.array_equal(npy(got.accepted_soa), ref["accepted"]), cls.__name__
        assert rel_err(npy(got.get_samples().permute(1, 0, 2)), ref["sample"]) < 1e-10


def test_runtime_shape_path_agrees_with_the_specialisation(monkeypatch):
    m = make_model("2321", "f64", 1.7)
    ds = dataset("2321", "f64")
    th = torch.randn(50, 20, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    a, ga = m.upto_grad_log_target_batch(th, ds.x, ds.y)
    monkeypatch.setenv("EEYORE_B200_FORCE_GENERIC", "1")
    m2 = make_model("2321", "f64", 1.7)
    b, gb = m2.upto_grad_log_target_batch(th, ds.x, ds.y)
    from eeyore_b200 import _native as nv
    assert nv.lib().eeyore_b200_mlp_is_specialised(m.handle()) == 1 and nv.lib().eeyore_b200_mlp_is_specialised(m2.handle()) == 0
    assert torch.allclose(a, b, rtol=1e-13, atol=0) and torch.allclose(ga, gb, rtol=1e-11, atol=1e-13)


def test_posterior_predictive_and_logistic_regression():
    """BayesianModel.predictive_posterior(_from_dataset) (bayesian_model.py:58-67,
    integrators/mcintegrator.py:16-63) as one batched launch, and LogisticRegression (models/logistic_regression.py)."""
    from eeyore_b200.constants import loss_functions
    from eeyore_b200.models import LogisticRegression
    from eeyore_b200.models.logistic_regression import Hyperparameters as LRH
    from oracle.mlp import MLPSpec
    m = make_model("433", "f64", 1.0)
    ds = dataset("433", "f64")
    rng = np.random.default_rng(0)
    S = 50
    th = rng.normal(size=(S, 27)) * 0.3
    th[7, 3] = np.nan                                                   # a NaN sample is dropped, mcintegrator.py:24-25
    val, dropped = m.predictive_posterior(torch.from_numpy(th), ds.x[:5], ds.y[:5])
    ll = oracle.log_lik(spec_of("433"), th, npy(ds.x[:5]), npy(ds.y[:5]))
    assert dropped == 1 and abs(val.item() - np.nanmean(np.exp(ll))) < 1e-12 * np.nanmean(np.exp(ll))
    integ, idx, nd = m.predictive_posterior_from_dataset(list(torch.from_numpy(th)), ds, 12, shuffle=False)
    for k in range(12):
        llk = oracle.log_lik(spec_of("433"), th, npy(ds.x[k:k + 1]), npy(ds.y[k:k + 1]))
        assert abs(integ[k].item() - np.nanmean(np.exp(llk))) < 1e-11 * np.nanmean(np.exp(llk))
    assert idx.tolist() == list(range(12)) and nd.tolist() == [1] * 12
    # logistic regression = one dense layer with a sigmoid head
    lr = LogisticRegression(loss=loss_functions["binary_classification"], hparams=LRH(input_size=6, output_size=1))
    assert lr.num_params() == 7
    x = rng.normal(size=(40, 6)); y = (rng.uniform(size=(40, 1)) < 0.5).astype(np.float64)
    thl = rng.normal(size=(9, 7))
    lt, g = lr.upto_grad_log_target_batch(torch.from_numpy(thl), torch.from_numpy(x), torch.from_numpy(y))
    spec = MLPSpec([6, 1], loss="binary_classification", bias=[True], activations=["sigmoid"]) if False else None
    # closed form: ll = sum y log s + (1-y) log(1-s), s = sigmoid(x w + b); prior N(0, 1)
    w_, b_ = thl[:, :6], thl[:, 6]
    s = 1 / (1 + np.exp(-(x @ w_.T + b_)))                              # [40, 9]
    ll_ref = (np.log(s) * y + np.log(1 - s) * (1 - y)).sum(0)
    lp_ref = (-0.5 * thl ** 2 - 0.5 * np.log(2 * np.pi)).sum(1)
    assert np.allclose(npy(lt), ll_ref + lp_ref, rtol=1e-11)
    g_ref = np.concatenate([((y - s).T @ x), (y - s).sum(0)[:, None]], axis=1) - thl
    assert rel_err(npy(g), g_ref) < 1e-11


def test_fast_and_general_forward_paths_agree_with_the_oracle_on_device():
    """GPU twin of tests/test_hostsim.py::test_fast_and_general_forward_paths_agree_with_the_oracle: chains of one warp
    straddle the bounds of the select-free fp64 fast path (|head pre-activation| = 36 and 708, hidden 708), so lanes of a
    warp diverge between the fast and the general row code; values, gradients and the NaN / -inf pattern must match the
    oracle on both sides (evaluation kernel, one lane and four lanes per chain)."""
    from test_hostsim import _boundary_thetas
    th = _boundary_thetas()
    m = make_model("2321", "f64", 3.0 ** 0.5)
    ds = dataset("2321", "f64")
    spec = spec_of("2321")
    loc, scale = np.zeros(20), np.full(20, 3.0 ** 0.5)
    with np.errstate(all="ignore"):
        lt_ref, g_ref = oracle.log_target_grad(spec, th, npy(ds.x), npy(ds.y), loc, scale)
    for lanes in (1, 4):
        lt, g = m.upto_grad_log_target_batch(torch.from_numpy(th), ds.x, ds.y, lanes=lanes)
        lt, g = npy(lt), npy(g)
        assert np.array_equal(np.isnan(lt), np.isnan(lt_ref)) and np.array_equal(np.isinf(lt), np.isinf(lt_ref))
        fin = np.isfinite(lt_ref)
        assert np.allclose(lt[fin], lt_ref[fin], rtol=1e-10, atol=0)
        for c in np.nonzero(fin)[0]:
            assert rel_err(g[c], g_ref[c]) < 1e-10, (lanes, c)
        assert np.isnan(g[np.isnan(lt_ref)]).all()


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_soft_labels_on_device(tag):
    """GPU twin of tests/test_hostsim.py::test_soft_labels_take_the_general_path (targets in (0, 1), stats/loss.py:2), plus a
    data set whose only soft label sits in the last row, so that the CTA-wide hard-label flag has to see every row."""
    from test_hostsim import _soft_label_case
    x, y, th = _soft_label_case()
    dt, tdt = NP_DTYPES[tag], T_DTYPES[tag]
    m = make_model("2321", tag, 2.0)
    loc, scale = np.zeros(20), np.full(20, 2.0)
    tol = RTOL[tag]
    up = lambda a: a.astype(dt).astype(np.float64)        # the oracle sees the inputs as rounded to the kernel's dtype
    for yy in (y, np.array([[0.0], [1.0], [1.0], [0.0], [1.0], [0.0], [0.25]])):
        lt_ref, g_ref = oracle.log_target_grad(spec_of("2321"), up(th), up(x), up(yy), loc, scale)
        for lanes in (1, 4, 32):
            lt, g = m.upto_grad_log_target_batch(torch.from_numpy(th.astype(dt)), torch.from_numpy(x.astype(dt)),
                                                 torch.from_numpy(yy.astype(dt)), lanes=lanes)
            assert np.allclose(npy(lt), lt_ref, rtol=tol, atol=0), lanes
            for c in range(th.shape[0]):
                assert rel_err(npy(g)[c], g_ref[c]) < tol, (lanes, c)
