"""Pins the numpy oracle against golden vectors produced by the unmodified reference
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

import oracle
from helpers import ARCHS, NP_DTYPES, PRIOR_SCALES, RTOL, data_of, load, rel_err, spec_of


@pytest.mark.parametrize("arch", list(ARCHS))
@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("pst", list(PRIOR_SCALES))
@pytest.mark.parametrize("tt,temp", [("", None), ("_T07", 0.7)])
def test_log_target_and_grad(arch, tag, pst, tt, temp):
    mg = load("model_goldens")
    dt = NP_DTYPES[tag]
    spec = spec_of(arch)
    x, y = data_of(arch, dt, mg)
    key = f"{arch}_{tag}_{pst}{tt}"
    theta = mg[key + "_theta"].astype(dt)
    P = theta.shape[1]
    loc, scale = np.zeros(P, dt), np.full(P, PRIOR_SCALES[pst], dt)
    lt, g = oracle.log_target_grad(spec, theta, x, y, loc, scale, temp)
    ll = oracle.log_lik(spec, theta, x, y, temp)
    lp = oracle.log_prior(theta, loc, scale, temp)
    tol = RTOL[tag] if tag == "f64" else 2e-5
    assert np.allclose(lt, mg[key + "_lt"], rtol=tol, atol=0)
    assert np.allclose(ll, mg[key + "_ll"], rtol=tol, atol=0)
    assert np.allclose(lp, mg[key + "_lp"], rtol=tol, atol=0)
    for c in range(theta.shape[0]):
        assert rel_err(g[c], mg[key + "_grad"][c]) < tol, (c, g[c], mg[key + "_grad"][c])


def test_reference_known_answers():
    """SURVEY.md section 4 known-answer table (values computed from the reference)."""
    mg = load("model_goldens")
    x, y = data_of("221", np.float64, mg)
    th = mg["221_f64_p1_theta"][:1]
    spec = spec_of("221")
    assert abs(oracle.log_lik(spec, th, x, y)[0] - (-16.08587869723768)) < 1e-12
    lt, g = oracle.log_target_grad(spec, th, x, y, np.zeros(9), np.ones(9))
    assert abs(lt[0] - (-122.68812549607972)) < 1e-11
    lt100, g100 = oracle.log_target_grad(spec, th, x, y, np.zeros(9), 100 * np.ones(9))
    assert abs(lt100[0] - (-65.81269034997256)) < 1e-11
    assert abs(g100[0, 0] - (-3.11250194173020467e-01)) < 1e-14
    xi, yi = data_of("433", np.float64, mg)
    th = mg["433_f64_p1_theta"][:1]
    assert abs(oracle.log_lik(spec_of("433"), th, xi, yi)[0] - (-176.25449918882558)) < 1e-11
    th = mg["4323_f64_p1_theta"][:1]
    assert abs(oracle.log_lik(spec_of("4323"), th, xi, yi)[0] - (-187.79398747047628)) < 1e-11
    th = mg["2321_f64_p1_theta"][:1]
    assert abs(oracle.log_lik(spec_of("2321"), th, x, y)[0] - (-7.39132872690876)) < 1e-12


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_saturation_is_nan(tag):
    """p saturating to exactly 1.0 gives NaN target and gradient, as in the reference."""
    mg = load("model_goldens")
    dt = NP_DTYPES[tag]
    x, y = data_of("221", dt, mg)
    th = mg[f"sat_{tag}_theta"].astype(dt)[None]
    assert np.isnan(mg[f"sat_{tag}_lt"]) and np.isnan(mg[f"sat_{tag}_grad"]).all()
    lt, g = oracle.log_target_grad(spec_of("221"), th, x, y, np.zeros(9, dt), np.ones(9, dt))
    assert np.isnan(lt[0]) and np.isnan(g).all()


def _check_run(name, runner, tag="f64", strict=True, **kw):
    gd = load(name)
    dt = NP_DTYPES[tag]
    arch = name.split("_")[1].replace("xor", "").replace("iris", "")
    spec = spec_of(arch)
    x, y = data_of(arch, dt)
    P = gd["theta0"].shape[0]
    loc, scale = np.zeros(P, dt), np.full(P, float(gd["prior_scale"]), dt)
    z = gd["z"].astype(dt)[:, None, :]
    u = gd["u"].astype(dt)[:, None]
    out = runner(spec, x, y, loc, scale, gd["theta0"].astype(dt)[None], z, u, n_burnin=int(gd["n_burnin"]), **kw)
    acc = out["accepted"][:, 0]
    if strict:
        assert np.array_equal(acc, gd["accepted"]), f"accept decisions differ at {np.nonzero(acc != gd['accepted'])[0][:5]}"
        tol = RTOL[tag] if tag == "f64" else 1e-3
        assert rel_err(out["sample"][:, 0], gd["samples"]) < tol
        assert rel_err(out["target_val"][:, 0], gd["target_vals"]) < tol
        if "grad_vals" in gd.files:
            assert rel_err(out["grad_val"][:, 0], gd["grad_vals"]) < max(tol, 1e-9)
    return out, gd


def test_mala_config1_trajectory():
    """BASELINE.json configs[0]: 2-2-1 XOR, 1100 iterations, 110 burn-in, fp64 -- identical accept vector."""
    out, gd = _check_run("mala_xor221_f64", oracle.mala_run, step=float(load("mala_xor221_f64")["step"]))
    assert out["sample"].shape == (990, 1, 9)
    ess = oracle.multi_ess(out["sample"][:, 0])
    assert abs(ess - float(gd["multi_ess"])) / float(gd["multi_ess"]) < 1e-6


def test_mala_iris_f64():
    _check_run("mala_iris433_f64", oracle.mala_run, step=0.003)


def test_mala_iris_f32_prefix():
    """fp32 trajectories are compared over the first iterations only (rounding differences between summation
    orders grow along a chain); accept decisions must agree on that prefix."""
    gd = load("mala_iris433_f32")
    dt = np.float32
    spec = spec_of("433")
    x, y = data_of("433", dt)
    P = 27
    loc, scale = np.zeros(P, dt), np.full(P, float(gd["prior_scale"]), dt)
    out = oracle.mala_run(spec, x, y, loc, scale, gd["theta0"][None], gd["z"][:, None, :], gd["u"][:, None],
                          step=0.003, n_burnin=0)
    nb = int(gd["n_burnin"])
    k = 10
    assert rel_err(out["sample"][nb:nb + k, 0], gd["samples"][:k]) < 1e-3


@pytest.mark.parametrize("name,step,L", [("hmc_xor2321_f64", 0.3, 10), ("hmc_xor221_f64", 0.9, 7),
                                         ("hmc_iris433_f64", 0.04, 10), ("hmc_xor2321_f64_s09", 0.9, 10)])
def test_hmc_trajectories(name, step, L):
    gd = load(name)
    assert float(gd["step"]) == step and int(gd["num_steps"]) == L
    _check_run(name, oracle.hmc_run, step=step, num_steps=L)


def test_mh_trajectories():
    _check_run("mh_xor221_f64", oracle.mh_run)
    _check_run("mh_xor2321_f64_nonsym", oracle.mh_run, symmetric=False, prop_scale=0.4)


def test_stats_goldens():
    gd = load("stats_goldens")
    x = gd["chains"]
    for i in range(4):
        assert rel_err(oracle.cov(x[i]), gd["cov"][i]) < 1e-12
        assert rel_err(oracle.inse_mc_cov(x[i]), gd["inse"][i]) < 1e-10
        assert abs(oracle.multi_ess(x[i]) - gd["multi_ess"][i]) / gd["multi_ess"][i] < 1e-10
    # SURVEY.md section 4 known answers
    assert abs(gd["multi_ess"][0] - 564.6937234344964) < 1e-9
    assert abs(oracle.cov(x[0])[0, 0] - 6.895458038624478) < 1e-12


def test_acf_against_third_party_implementations():
    """ACF is absent from the reference (kanga, un-vendored): the restatement is checked against scipy.signal.correlate and
    numpy.correlate evaluated over the reference's chain fixtures (oracle/make_golden.py acf_goldens)."""
    gd, ga = load("stats_goldens"), load("acf_goldens")
    x, k = gd["chains"], int(ga["max_lag"])
    for i in range(4):
        assert np.max(np.abs(oracle.acf(x[i], k) - ga["acf"][i])) < 1e-13
        assert np.max(np.abs(oracle.acf(x[i], k) - ga["acf_fft"][i])) < 1e-13


def test_inse_not_enough_samples():
    x = np.array([[0.0, 1.0], [1.0, 0.0], [0.5, 0.5]])
    with pytest.raises(RuntimeError, match="Not enough samples"):
        oracle.inse_mc_cov(x)


@pytest.mark.parametrize("name,l,e0,eub", [("hmcda_xor2321_f64", 0.6, 0.05, None), ("hmcda_iris433_f64", 0.15, 0.01, 0.05)])
def test_hmc_with_dual_averaging_tuner(name, l, e0, eub):
    """HMC + HMCDATuner (eeyore/tuners/hmcda_tuner.py, hmc.py:158-163): step size and trajectory length adapt in burn-in."""
    tuner = oracle.DATuner(l=l, e0=e0, n_chains=1, eub=eub)
    out, gd = _check_run(name, oracle.hmc_run, step=e0, num_steps=1, tuner=tuner)
    assert abs(out["final"]["step"][0] - float(gd["final_step"])) < 1e-12 * float(gd["final_step"])
    assert int(out["final"]["num_steps"][0]) == int(gd["final_num_steps"])


@pytest.mark.parametrize("name", ["pp_221_mix", "pp_2321_mala"])
def test_power_posterior_matches_the_reference(name):
    """Tempered MH / MALA chains with neighbour swaps (power_posterior_sampler.py) fed the reference's own noise,
    categorical draws and accept uniforms: every level's saved chain within 1e-10."""
    from helpers import pp_setup
    from oracle.power_posterior import power_posterior_run
    gd, spec, x, y, kinds, kwargs = pp_setup(name)
    P = gd["theta0"].shape[0]
    s3 = 3.0 ** 0.5
    r = power_posterior_run(spec, x, y, np.zeros(P), np.full(P, s3), gd["theta0"][None], kinds, kwargs,
                            gd["z"][:, :, None, :], gd["u"][:, :, None], gd["j_tape"][:, :, None], gd["u_between"][:, :, None],
                            temperatures=list(gd["temperatures"]), between_step=int(gd["between_step"]), b=float(gd["b"]),
                            n_burnin=int(gd["n_burnin"]))
    assert r["sample"].shape[:2] == gd["samples"].shape[:2]
    assert rel_err(r["sample"][:, :, 0], gd["samples"]) < 1e-10
    assert np.allclose(r["target_val"][:, :, 0], gd["target_vals"], rtol=1e-10, atol=1e-12)
    assert rel_err(r["final_sample"][:, 0], gd["final_sample"]) < 1e-10
    assert 0 < r["swaps"].sum() < r["swaps"].size          # the run contains accepted and rejected swaps


@pytest.mark.parametrize("name", ["am_xor221_f64", "am_xor2321_f64", "ram_xor221_f64", "ram_xor2321_f64"])
def test_adaptive_metropolis_matches_the_reference(name):
    """AM (am.py:62-107) and RAM (ram.py:39-70) fed the reference's noise: identical accept vectors, states < 1e-10 ...
    up to the adaptation's own conditioning (the factor is re-derived from sums of outer products every iteration)."""
    gd = load(name)
    arch = "221" if "221" in name else "2321"
    spec = spec_of(arch)
    x, y = data_of(arch, np.float64)
    P = gd["theta0"].shape[0]
    s3 = 3.0 ** 0.5
    if name.startswith("am"):
        r = oracle.am_run(spec, x, y, np.zeros(P), np.full(P, s3), gd["theta0"][None], gd["z"][:, None], gd["u"][:, :, None],
                          n_burnin=int(gd["n_burnin"]), l=float(gd["l"]), b=float(gd["b"]), c=float(gd["c"]), t0=int(gd["t0"]))
        factor = r["final"]["cov"][0]
    else:
        r = oracle.ram_run(spec, x, y, np.zeros(P), np.full(P, s3), gd["theta0"][None], gd["z"][:, None], gd["u"][:, None],
                           n_burnin=int(gd["n_burnin"]), a=float(gd["a"]), g=float(gd["g"]))
        factor = r["final"]["chol_cov"][0]
    assert np.array_equal(r["accepted"][:, 0], gd["accepted"])
    assert rel_err(r["sample"][:, 0], gd["samples"]) < 1e-9
    assert np.allclose(r["target_val"][:, 0], gd["target_vals"], rtol=1e-9, atol=1e-11)
    assert rel_err(factor, gd["final_factor"]) < 1e-7


@pytest.mark.parametrize("name,arch", [("smmala_xor2321_f64", "2321"), ("smmala_nxor2321_f64", "2321"),
                                       ("smmala_xor221_f64", "221")])
def test_smmala_restatement_matches_the_reference_pieces_run(name, arch):
    """oracle.smmala_run against a run assembled from the reference's own pieces (make_golden.py:smmala_goldens: reference
    MLP + autograd row derivatives for the metric, is_pos_def, torch.linalg.cholesky, MultivariateNormalKernel.log_prob)."""
    gd = load(name)
    spec = spec_of(arch)
    P = spec.num_params
    ref = oracle.smmala_run(spec, gd["x"], gd["y"], np.zeros(P), np.full(P, float(gd["prior_scale"])), gd["theta0"][None],
                            gd["z"][:, None, :], gd["u"][:, None], float(gd["step"]), n_burnin=int(gd["n_burnin"]))
    assert np.array_equal(ref["accepted"][:, 0], gd["accepted"])
    assert rel_err(ref["sample"][:, 0], gd["samples"]) < 1e-12
    assert rel_err(ref["target_val"][:, 0], gd["target_vals"]) < 1e-12
    assert rel_err(ref["grad_val"][:, 0], gd["grad_vals"]) < 1e-11


def test_fp32_fixtures_evaluated_in_fp64_by_the_reference():
    """model_goldens_f32ref.npz (the reference in fp64 at the fp32 fixtures): the oracle reproduces it at 1e-12, so the
    1e-5 bar of the fp32 kernels can be measured against fp64 truth instead of torch's own fp32 rounding."""
    mg, ref = load("model_goldens"), load("model_goldens_f32ref")
    for arch in ARCHS:
        x, y = data_of(arch, np.float32, mg)
        spec = spec_of(arch)
        P = spec.num_params
        for pst, ps in PRIOR_SCALES.items():
            for tt, temp in (("", None), ("_T07", 0.7)):
                key = f"{arch}_f32_{pst}{tt}"
                lt, g = oracle.log_target_grad(spec, mg[key + "_theta"].astype(np.float64), x.astype(np.float64),
                                               y.astype(np.float64), np.zeros(P), np.full(P, float(np.float32(ps))), temp)
                assert np.allclose(lt, ref[key + "_lt64"], rtol=1e-12, atol=0), key
                assert rel_err(g, ref[key + "_grad64"]) < 1e-12, key
                # and torch's fp32 evaluation of the same fixtures is itself within the fp32 bar of that truth
                assert rel_err(mg[key + "_grad"], ref[key + "_grad64"]) < 2e-5, key
