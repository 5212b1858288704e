"""CPU pre-flight of the DEVICE code: tests/hostsim compiles the __host__ __device__ per-chain functions of
eeyore_b200/csrc (mlp_static.cuh, samplers.cuh, philox.cuh) with g++ and this file checks them against the golden
vectors of the reference and against the oracle.  It is not a product path (nothing in eeyore_b200 loads it)."""
import ctypes as C

import numpy as np
import pytest

import oracle
from eeyore_b200._native import RunParams, F32, F64
from helpers import ARCHS, NP_DTYPES, PRIOR_SCALES, RTOL, data_of, load, rel_err, spec_of


@pytest.fixture(scope="module")
def sim():
    import __graft_entry__ as g
    lib = C.CDLL(str(g.build_hostsim()))
    lib.hostsim_eval.argtypes = [C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 2 + \
                                [C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    lib.hostsim_run.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(RunParams)]
    lib.hostsim_philox.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    return lib


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def dt_id(tag):
    return F64 if tag == "f64" else F32


def sim_eval(sim, arch, tag, theta, x, y, loc, scale, temp=None, grad=True):
    dt = NP_DTYPES[tag]
    theta = np.ascontiguousarray(theta, dt)
    x = np.ascontiguousarray(x, dt); y = np.ascontiguousarray(y, dt)
    lt = np.empty(theta.shape[0], dt)
    g = np.empty_like(theta) if grad else None
    rc = sim.hostsim_eval(int(arch), dt_id(tag), theta.shape[0], P(theta), P(x), P(y), x.shape[0], P(loc), P(scale),
                          0 if temp is None else 1, 0.0 if temp is None else temp, P(lt), P(g) if grad else None)
    assert rc == 0
    return lt, g


@pytest.mark.parametrize("arch", list(ARCHS))
@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("pst", ["p1", "psqrt3"])
@pytest.mark.parametrize("tt,temp", [("", None), ("_T07", 0.7)])
def test_device_eval_matches_reference_goldens(sim, arch, tag, pst, tt, temp):
    mg = load("model_goldens")
    dt = NP_DTYPES[tag]
    x, y = data_of(arch, dt, mg)
    key = f"{arch}_{tag}_{pst}{tt}"
    theta = mg[key + "_theta"].astype(dt)
    n = theta.shape[1]
    loc, scale = np.zeros(n, dt), np.full(n, PRIOR_SCALES[pst], dt)
    lt, g = sim_eval(sim, arch, tag, theta, x, y, loc, scale, temp)
    tol = RTOL[tag]
    if tag == "f64":
        want_lt, want_g = mg[key + "_lt"], mg[key + "_grad"]
    else:                                 # fp64 truth at the fp32 fixtures (the reference run in fp64)
        ref64 = load("model_goldens_f32ref")
        want_lt, want_g = ref64[key + "_lt64"], ref64[key + "_grad64"]
    assert np.allclose(lt, want_lt, rtol=tol, atol=0)
    for c in range(theta.shape[0]):
        assert rel_err(g[c], want_g[c]) < tol


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_device_saturation_nan(sim, tag):
    mg = load("model_goldens")
    dt = NP_DTYPES[tag]
    x, y = data_of("221", dt, mg)
    th = mg[f"sat_{tag}_theta"].astype(dt)[None]
    lt, g = sim_eval(sim, "221", tag, th, x, y, np.zeros(9, dt), np.ones(9, dt))
    assert np.isnan(lt[0]) and np.isnan(g).all()


def run_params(dt, theta, lt, g, x, y, loc, scale, z, u, n_burnin, step, L=1, symmetric=1, seed=0, tape=True):
    C_, n = theta.shape
    T = z.shape[0] if tape else int(u)
    ns = T - n_burnin
    out = dict(sample=np.zeros((ns, C_, n), dt), target=np.zeros((ns, C_), dt), grad=np.zeros((ns, C_, n), dt),
               acc=np.zeros((ns, C_), np.uint8), count=np.zeros(C_, np.uint32))
    p = RunParams()
    p.n_chains, p.n_iters, p.n_burnin, p.thin = C_, T, n_burnin, 1
    p.step, p.num_steps, p.symmetric = step, L, symmetric
    p.rng_mode = 1 if tape else 0
    p.seed = seed
    if tape:
        p.z_tape, p.u_tape = z.ctypes.data, u.ctypes.data
    p.x, p.y, p.n_rows = x.ctypes.data, y.ctypes.data, x.shape[0]
    p.prior_loc, p.prior_scale = loc.ctypes.data, scale.ctypes.data
    p.theta, p.target = theta.ctypes.data, lt.ctypes.data
    p.grad = g.ctypes.data if g is not None else None
    p.out_samples, p.ss_iter, p.ss_chain, p.ss_param = out["sample"].ctypes.data, C_ * n, n, 1
    p.out_target, p.out_grad, p.out_accepted = out["target"].ctypes.data, out["grad"].ctypes.data, out["acc"].ctypes.data
    p.accept_count = out["count"].ctypes.data
    return p, out


KINDS = {"mh": 0, "mala": 1, "hmc": 2}


@pytest.mark.parametrize("name,kind,kw", [
    ("mala_xor221_f64", "mala", dict(step=1.74)),
    ("mala_iris433_f64", "mala", dict(step=0.003)),
    ("hmc_xor2321_f64", "hmc", dict(step=0.3, L=10)),
    ("hmc_xor2321_f64_s09", "hmc", dict(step=0.9, L=10)),
    ("hmc_xor221_f64", "hmc", dict(step=0.9, L=7)),
    ("hmc_iris433_f64", "hmc", dict(step=0.04, L=10)),
    ("mh_xor221_f64", "mh", dict(step=1.0)),
    ("mh_xor2321_f64_nonsym", "mh", dict(step=0.4, symmetric=0)),
])
def test_device_samplers_reproduce_reference_runs(sim, name, kind, kw):
    """Fed the reference's proposal noise, the device sampler code reproduces its accept decisions and states."""
    gd = load(name)
    dt = np.float64
    arch = name.split("_")[1].replace("xor", "").replace("iris", "")
    x, y = data_of(arch, dt)
    x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
    n = gd["theta0"].shape[0]
    loc, scale = np.zeros(n, dt), np.full(n, float(gd["prior_scale"]), dt)
    theta = gd["theta0"].astype(dt)[None].copy()
    lt, g = sim_eval(sim, arch, "f64", theta, x, y, loc, scale)
    z = np.ascontiguousarray(gd["z"].astype(dt)[:, None, :]); u = np.ascontiguousarray(gd["u"].astype(dt)[:, None])
    p, out = run_params(dt, theta, lt, g, x, y, loc, scale, z, u, int(gd["n_burnin"]), **kw)
    assert sim.hostsim_run(KINDS[kind], int(arch), F64, C.byref(p)) == 0
    assert np.array_equal(out["acc"][:, 0], gd["accepted"])
    assert rel_err(out["sample"][:, 0], gd["samples"]) < 1e-10
    assert rel_err(out["target"][:, 0], gd["target_vals"]) < 1e-10
    if kind != "mh":
        assert rel_err(out["grad"][:, 0], gd["grad_vals"]) < 1e-9
    assert rel_err(theta[0], gd["final_sample"]) < 1e-10
    total_acc = int(out["count"][0])
    assert total_acc >= int(gd["accepted"].sum())


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_device_philox_matches_oracle_stream(sim, tag):
    dt = NP_DTYPES[tag]
    Cn, Pn, seed, it, c0 = 37, 27, 0x1234567890ABCDEF, 12345, 1000
    z = np.empty((Cn, Pn), dt); u = np.empty(Cn, dt)
    assert sim.hostsim_philox(dt_id(tag), Cn, Pn, seed, it, c0, P(z), P(u)) == 0
    zr = oracle.chain_normals(seed, np.arange(c0, c0 + Cn), it, Pn, dt)
    ur = oracle.chain_uniforms(seed, np.arange(c0, c0 + Cn), it, dt)
    tol = 1e-12 if tag == "f64" else 2e-5
    assert np.max(np.abs(z - zr)) < tol * 10
    assert np.array_equal(u, ur)
    assert 0 < u.min() and u.max() < 1


def test_philox_known_answer():
    """Random123 known-answer vectors for philox4x32-10."""
    r = oracle.philox4x32_10([np.uint32(0)] * 4, (0, 0))
    assert [int(v) for v in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = oracle.philox4x32_10([np.uint32(0xffffffff)] * 4, (0xffffffff, 0xffffffff))
    assert [int(v) for v in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = oracle.philox4x32_10([np.uint32(0x243f6a88), np.uint32(0x85a308d3), np.uint32(0x13198a2e), np.uint32(0x03707344)],
                             (0xa4093822, 0x299f31d0))
    assert [int(v) for v in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_device_philox_hmc_matches_oracle_with_same_stream(sim):
    """Philox mode end to end: device code draws its own noise; the oracle is fed the oracle-side replica of the stream."""
    dt = np.float64
    arch, n, Cn, T, L, step, seed = "2321", 20, 5, 12, 6, 0.4, 99
    x, y = data_of(arch, dt)
    x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
    loc, scale = np.zeros(n, dt), np.full(n, 3 ** 0.5, dt)
    rng = np.random.default_rng(5)
    theta0 = rng.normal(size=(Cn, n))
    theta = theta0.copy()
    lt, g = sim_eval(sim, arch, "f64", theta, x, y, loc, scale)
    p, out = run_params(dt, theta, lt, g, x, y, loc, scale, None, T, 0, step, L=L, seed=seed, tape=False)
    p.chain_offset, p.iter_offset = 7, 3
    assert sim.hostsim_run(2, int(arch), F64, C.byref(p)) == 0
    z = np.stack([oracle.chain_normals(seed, np.arange(7, 7 + Cn), 3 + t, n) for t in range(T)])
    u = np.stack([oracle.chain_uniforms(seed, np.arange(7, 7 + Cn), 3 + t) for t in range(T)])
    ref = oracle.hmc_run(spec_of(arch), x, y, loc, scale, theta0, z, u, step, L)
    assert np.array_equal(out["acc"], ref["accepted"])
    assert rel_err(out["sample"], ref["sample"]) < 1e-9


@pytest.mark.parametrize("name,arch,l,e0,eub", [("hmcda_xor2321_f64", "2321", 0.6, 0.05, None),
                                                 ("hmcda_iris433_f64", "433", 0.15, 0.01, 0.05)])
def test_device_hmc_dual_averaging_tuner_reproduces_reference(sim, name, arch, l, e0, eub):
    """HMC + HMCDATuner of the reference (hmc.py:158-163, tuners/hmcda_tuner.py): the device recurrence reproduces the
    adapted step sizes, trajectory lengths and therefore the accept decisions of the reference run."""
    import math
    gd = load(name)
    dt = np.float64
    x, y = data_of(arch, dt)
    x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
    n = gd["theta0"].shape[0]
    loc, scale = np.zeros(n, dt), np.full(n, float(gd["prior_scale"]), dt)
    theta = gd["theta0"].astype(dt)[None].copy()
    lt, g = sim_eval(sim, arch, "f64", theta, x, y, loc, scale)
    z = np.ascontiguousarray(gd["z"].astype(dt)[:, None, :]); u = np.ascontiguousarray(gd["u"].astype(dt)[:, None])
    nb = int(gd["n_burnin"])
    p, out = run_params(dt, theta, lt, g, x, y, loc, scale, z, u, nb, e0, L=max(1, round(l / e0)))
    state = np.array([[0.0], [0.0], [e0], [float(max(1, round(l / e0)))]])
    p.tuner_l, p.tuner_d, p.tuner_m = l, 0.65, math.log(10 * e0)
    p.tuner_has_eub, p.tuner_logeub = (0, 0.0) if eub is None else (1, math.log(eub))
    p.tuner_iter0, p.tuner_burnin, p.tuner_state = 0, nb, state.ctypes.data
    assert sim.hostsim_run(2, int(arch), F64, C.byref(p)) == 0
    assert np.array_equal(out["acc"][:, 0], gd["accepted"])
    assert rel_err(out["sample"][:, 0], gd["samples"]) < 1e-9
    assert abs(state[2, 0] - float(gd["final_step"])) < 1e-11 * float(gd["final_step"])
    assert int(state[3, 0]) == int(gd["final_num_steps"])


GEN_CASES = [
    # dims, bias, activations (None = identity), loss
    ([3, 5, 1], [True, True], ["sigmoid", "sigmoid"], "binary_classification"),
    ([2, 3, 2, 1], [True, True, True], ["sigmoid", "sigmoid", "sigmoid"], "binary_classification"),
    ([4, 3, 3], [True, False], ["sigmoid", None], "multiclass_classification"),
    ([4, 6, 5, 4, 3], [True, False, True, True], ["sigmoid", None, "sigmoid", None], "multiclass_classification"),
    ([2, 4, 1], [False, False], [None, "sigmoid"], "binary_classification"),
    ([5, 7, 3, 2, 4, 1], [True] * 5, ["sigmoid"] * 5, "binary_classification"),
]


def gen_problem(dims, loss, n_rows, seed):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n_rows, dims[0]))
    k = dims[-1]
    y = rng.integers(0, 2, size=(n_rows, 1)).astype(np.float64) if k == 1 else np.eye(k)[rng.integers(0, k, size=n_rows)]
    return x, y


@pytest.mark.parametrize("dims,bias,acts,loss", GEN_CASES)
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_runtime_shape_eval_matches_oracle(sim, dims, bias, acts, loss, tag):
    """Arbitrary dims / bias flags / activations (mlp.Hyperparameters, eeyore/models/mlp.py:9-19) through the runtime-shape
    device code."""
    from oracle.mlp import MLPSpec
    sim.hostsim_gen_eval.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int,
                                     C.c_int64] + [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 2 + [C.c_int, C.c_double,
                                                                                                      C.c_void_p, C.c_void_p]
    dt = NP_DTYPES[tag]
    spec = MLPSpec(dims, loss=loss, bias=bias, activations=acts)
    x, y = gen_problem(dims, loss, 23, sum(dims))
    rng = np.random.default_rng(1)
    theta = (rng.normal(size=(9, spec.num_params)) * 0.7).astype(dt)
    loc = (rng.normal(size=spec.num_params) * 0.1).astype(dt)
    scale = (0.5 + rng.uniform(size=spec.num_params)).astype(dt)
    nl = len(dims) - 1
    lt = np.empty(9, dt); g = np.empty_like(theta)
    xs, ys = np.ascontiguousarray(x, dt), np.ascontiguousarray(y, dt)
    P_ = sim.hostsim_gen_eval(nl, (C.c_int * (nl + 1))(*dims), (C.c_int * nl)(*[int(b) for b in bias]),
                              (C.c_int * nl)(*[1 if a else 0 for a in acts]), 0 if loss == "binary_classification" else 1,
                              dt_id(tag), 9, P(theta), P(xs), P(ys), 23, P(loc), P(scale), 1, 0.9, P(lt), P(g))
    assert P_ == spec.num_params
    lt_ref, g_ref = oracle.log_target_grad(spec, theta.astype(np.float64), xs.astype(np.float64), ys.astype(np.float64),
                                           loc.astype(np.float64), scale.astype(np.float64), 0.9)
    tol = 1e-11 if tag == "f64" else 1e-5
    assert np.allclose(lt, lt_ref, rtol=tol, atol=0)
    for c in range(9):
        assert rel_err(g[c], g_ref[c]) < tol


@pytest.mark.parametrize("kind", ["mh", "mala", "hmc"])
def test_runtime_shape_samplers_match_oracle(sim, kind):
    from oracle.mlp import MLPSpec
    sim.hostsim_gen_run.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int,
                                    C.POINTER(RunParams)]
    dims, bias, acts, loss = GEN_CASES[3]
    spec = MLPSpec(dims, loss=loss, bias=bias, activations=acts)
    dt = np.float64
    x, y = gen_problem(dims, loss, 17, 3)
    n, Cn, T = spec.num_params, 4, 9
    rng = np.random.default_rng(2)
    theta0 = rng.normal(size=(Cn, n)) * 0.4
    z, u = rng.normal(size=(T, Cn, n)), rng.uniform(size=(T, Cn))
    loc, scale = np.zeros(n), np.full(n, 1.3)
    kw = dict(mh=dict(step=0.05), mala=dict(step=0.03), hmc=dict(step=0.06, L=4))[kind]
    if kind == "mh":
        ref = oracle.mh_run(spec, x, y, loc, scale, theta0, z, u, n_burnin=2, prop_scale=0.05)
    elif kind == "mala":
        ref = oracle.mala_run(spec, x, y, loc, scale, theta0, z, u, 0.03, n_burnin=2)
    else:
        ref = oracle.hmc_run(spec, x, y, loc, scale, theta0, z, u, 0.06, 4, n_burnin=2)
    theta = theta0.copy()
    lt_, g_ = oracle.log_target_grad(spec, theta, x, y, loc, scale)
    lt, g = np.ascontiguousarray(lt_), np.ascontiguousarray(g_)
    xs, ys = np.ascontiguousarray(x), np.ascontiguousarray(y)
    p, out = run_params(dt, theta, lt, g, xs, ys, loc, scale, np.ascontiguousarray(z), np.ascontiguousarray(u), 2, **kw)
    nl = len(dims) - 1
    assert sim.hostsim_gen_run(KINDS[kind], nl, (C.c_int * (nl + 1))(*dims), (C.c_int * nl)(*[int(b) for b in bias]),
                               (C.c_int * nl)(*[1 if a else 0 for a in acts]), 1, F64, C.byref(p)) == 0
    assert ref["accepted"].mean() > 0
    assert np.array_equal(out["acc"], ref["accepted"])
    assert rel_err(out["sample"], ref["sample"]) < 1e-10


def _boundary_thetas():
    """2-3-2-1 parameter vectors whose head pre-activation sweeps across the bounds of the fp64 fast path (|a| = 36: p may
    round to exactly 1; |a| = 708: the exponent arithmetic of the fast exp would wrap) and whose hidden pre-activations
    cross 708, next to ordinary ones -- so that within one batch of chains (and, on the GPU, within one warp) some rows take
    the fast path and some the general one."""
    rng = np.random.default_rng(5)
    th = rng.normal(size=(96, 20)) * 0.7
    heads = [0.0, 20.0, 35.9, 36.1, 36.8, 40.0, 700.0, 709.0, 720.0, -20.0, -35.9, -36.1, -40.0, -700.0, -707.9, -708.1, -745.0,
             -800.0, 1e6, -1e6]
    for k, b in enumerate(heads):
        th[k, 17:19] = 0.0          # W2 = 0: the head pre-activation is the bias alone
        th[k, 19] = b
    for k, b in enumerate([30.0, 700.0, 707.0, 709.0, 5000.0, -700.0, -709.0, -5000.0]):
        th[32 + k, 6] = b           # b0[0]: a hidden unit of the first layer saturates
        th[40 + k, 15] = b          # b1[0]: one of the second layer
    return th


def test_fast_and_general_forward_paths_agree_with_the_oracle(sim):
    """The fp64 row code has a select-free fast path and a general path that reproduces the reference's saturation / NaN
    semantics (mlp_static.cuh: accumulate_row, accumulate_rows_fast).  Values, gradients and the NaN pattern must match the
    oracle on both sides of every switch point; rows are also permuted so that the two-row batches pair differently."""
    mg = load("model_goldens")
    x, y = data_of("2321", np.float64, mg)
    th = _boundary_thetas()
    loc, scale = np.zeros(20), np.full(20, 3.0 ** 0.5)
    spec = spec_of("2321")
    with np.errstate(all="ignore"):
        lt_ref, g_ref = oracle.log_target_grad(spec, th, x, y, loc, scale)
    for perm in ([0, 1, 2, 3], [3, 1, 0, 2], [2, 3, 1, 0]):
        lt, g = sim_eval(sim, "2321", "f64", th, x[perm], y[perm], loc, scale)
        assert np.array_equal(np.isnan(lt), np.isnan(lt_ref))
        assert np.array_equal(np.isinf(lt), np.isinf(lt_ref))
        fin = np.isfinite(lt_ref)
        assert fin.sum() > 60 and (~fin).sum() > 8
        assert np.allclose(lt[fin], lt_ref[fin], rtol=1e-10, atol=0)
        for c in np.nonzero(fin)[0]:
            assert rel_err(g[c], g_ref[c]) < 1e-10, c
        assert np.isnan(g[np.isnan(lt_ref)]).all()


def _soft_label_case():
    rng = np.random.default_rng(11)
    x = rng.normal(size=(7, 2))
    y = np.array([[0.0], [1.0], [0.3], [1.0], [0.75], [0.0], [0.5]])   # stats/loss.py:2 takes any float target
    th = rng.normal(size=(40, 20)) * 0.9
    return x, y, th


def test_soft_labels_take_the_general_path(sim):
    """Targets other than 0 / 1 (the reference's naive cross-entropy accepts them): the CTA-wide hard-label flag is off, every
    row goes through the general code with both logs."""
    x, y, th = _soft_label_case()
    loc, scale = np.zeros(20), np.full(20, 2.0)
    lt_ref, g_ref = oracle.log_target_grad(spec_of("2321"), th, x, y, loc, scale)
    lt, g = sim_eval(sim, "2321", "f64", th, x, y, loc, scale)
    assert np.allclose(lt, lt_ref, rtol=1e-10, atol=0)
    for c in range(th.shape[0]):
        assert rel_err(g[c], g_ref[c]) < 1e-10
