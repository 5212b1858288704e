"""Generates tests/golden/*.npz by running the UNMODIFIED reference (read-only at /root/reference).

Run in the build container only (the reference does not exist on the GPU box):
    python oracle/make_golden.py
The committed .npz files are what tests read; this script documents how they were produced.

Noise tape: torch.randn / torch.rand / torch.normal are wrapped for the duration of a sampler run so that
every standard-normal vector and accept-uniform the reference consumes is recorded in call order.
torch.normal(loc, scale) is replaced by loc + scale * torch.randn(...) (bitwise identical on CPU for the
same generator state), so the recorded z reproduces the reference's proposals.
Fixture parameter vectors are the ones of the reference's own unit tests (cited below).
"""
import math
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT / "oracle" / "kanga_stub"))
sys.path.insert(0, str(REF))

from torch.distributions import Normal  # noqa: E402
from torch.utils.data import DataLoader  # noqa: E402

from eeyore.chains import ChainList  # noqa: E402
from eeyore.constants import loss_functions  # noqa: E402
from eeyore.datasets import XYDataset  # noqa: E402
from eeyore.models.mlp import MLP, Hyperparameters  # noqa: E402
from eeyore.samplers import HMC, MALA, MetropolisHastings  # noqa: E402
from eeyore.tuners import HMCDATuner  # noqa: E402
import eeyore.stats as st  # noqa: E402

OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)

# fixture thetas of the reference's tests
THETA_221 = [1.1, -2.9, -0.4, 0.8, 4.3, 9.2, 4.44, -3.4, 7.2]            # tests/test_binary_classif_mlp221_log_lik.py:36
THETA_2321 = [1.1, -2.9, -0.4, 0.8, 4.3, 9.2, 4.44, -3.4, 7.2, 1.2, -2.3, 0.4, -5.4, -3.3, 2.8, 2.9, 7.7, -4.4, 2,
              6]                                                        # tests/test_binary_classif_mlp2321_log_lik.py:19-22 (first 20)
THETA_433 = [0.7735, 0.8161, 0.3910, 0.9622, 0.3748, 0.8711, 0.3315, 0.5473, 0.8820,
             0.0294, 0.9686, 0.8313, 0.6693, 0.8791, 0.6271, 0.8636, 0.3814, 0.0319,
             0.5148, 0.5086, 0.7428, 0.5464, 0.5278, 0.6127, 0.4499, 0.1538, 0.9291]   # tests/test_multiclass_classif_mlp433_log_lik.py:36-39
THETA_4323 = [0.2213, 0.5852, 0.1458, 0.5139, -0.1946, 0.0489, -0.1281, -0.7307,
              0.2176, 0.3274, -1.3060, 0.3253, -0.4248, 1.7403, 0.6219, 0.2652,
              -0.5310, -0.0291, 1.0262, -0.4920, 0.4391, -0.2450, 2.3145, -0.0788,
              1.1180, -1.2803, -0.4435, 0.5371, -0.2440, -0.3574, 0.4446, -0.3453]     # tests/test_multiclass_classif_mlp4323_log_lik.py:19-25

ARCHS = {
    "221": dict(dims=[2, 2, 1], loss="binary_classification", data="xor", theta=THETA_221),
    "2321": dict(dims=[2, 3, 2, 1], loss="binary_classification", data="xor", theta=THETA_2321),
    "433": dict(dims=[4, 3, 3], loss="multiclass_classification", data="iris", theta=THETA_433),
    "4323": dict(dims=[4, 3, 2, 3], loss="multiclass_classification", data="iris", theta=THETA_4323),
}


def load_data(name, dtype):
    if name == "xor":
        return XYDataset.from_eeyore("xor", dtype=dtype)
    return XYDataset.from_eeyore("iris", yndmin=1, yonehot=True, dtype=dtype)


def make_model(arch, dtype, prior_scale=None, temperature=None):
    a = ARCHS[arch]
    nl = len(a["dims"]) - 1
    last = torch.sigmoid if a["loss"] == "binary_classification" else None
    hp = Hyperparameters(dims=a["dims"], bias=nl * [True], activations=(nl - 1) * [torch.sigmoid] + [last])
    model = MLP(loss=loss_functions[a["loss"]], hparams=hp, dtype=dtype, temperature=temperature)
    if prior_scale is not None:
        P = model.num_params()
        model.prior = Normal(torch.zeros(P, dtype=dtype), prior_scale * torch.ones(P, dtype=dtype))
    return model


def npy(t):
    return t.detach().cpu().numpy()


# ---------------------------------------------------------------------------------------------------
def model_goldens():
    out = {}
    for dname in ("xor", "iris"):
        d = load_data(dname, torch.float64)
        out[f"{dname}_x"], out[f"{dname}_y"] = npy(d.x), npy(d.y)
    gen = torch.Generator().manual_seed(1234)
    for arch, a in ARCHS.items():
        for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            data = load_data(a["data"], dtype)
            for ps, pst in ((1.0, "p1"), (100.0, "p100"), (math.sqrt(3.0), "psqrt3")):
                for temp, tt in ((None, ""), (0.7, "_T07")):
                    model = make_model(arch, dtype, prior_scale=ps, temperature=temp)
                    P = model.num_params()
                    fixed = torch.tensor(a["theta"], dtype=dtype)
                    rnd = (torch.randn(12, P, generator=gen, dtype=torch.float64) * 1.5).to(dtype)
                    thetas = torch.cat([fixed[None], rnd])
                    lts, grads, lls, lps = [], [], [], []
                    for th in thetas:
                        lt, g = model.upto_grad_log_target(th.clone().detach(), data.x, data.y)
                        lts.append(lt.item()); grads.append(npy(g))
                        lls.append(model.log_lik(data.x, data.y).item()); lps.append(model.log_prior().item())
                    key = f"{arch}_{tag}_{pst}{tt}"
                    out[key + "_theta"] = npy(thetas)
                    out[key + "_lt"] = np.array(lts); out[key + "_grad"] = np.stack(grads)
                    out[key + "_ll"] = np.array(lls); out[key + "_lp"] = np.array(lps)
    # saturation semantics: output bias pushed until p == 1.0 exactly
    for dtype, tag, b in ((torch.float32, "f32", 20.0), (torch.float64, "f64", 40.0)):
        model = make_model("221", dtype, prior_scale=1.0)
        data = load_data("xor", dtype)
        th = torch.tensor(THETA_221, dtype=dtype); th[8] = b; th[6] = 0.0; th[7] = 0.0
        lt, g = model.upto_grad_log_target(th.clone().detach(), data.x, data.y)
        out[f"sat_{tag}_theta"] = npy(th); out[f"sat_{tag}_lt"] = np.array(lt.item()); out[f"sat_{tag}_grad"] = npy(g)
    np.savez_compressed(OUT / "model_goldens.npz", **out)
    print("model_goldens", len(out))


# ---------------------------------------------------------------------------------------------------
class Tape:
    """Records z (standard normals) and u (uniforms) in the reference's call order."""

    def __enter__(self):
        self.z, self.u = [], []
        self._randn, self._rand, self._normal = torch.randn, torch.rand, torch.normal

        def randn(*a, **k):
            v = self._randn(*a, **k); self.z.append(npy(v).copy()); return v

        def rand(*a, **k):
            v = self._rand(*a, **k); self.u.append(npy(v).copy()); return v

        def normal(loc, scale, *a, **k):
            assert not a and not k
            z = self._randn(loc.shape, dtype=loc.dtype); self.z.append(npy(z).copy())
            return loc + scale * z

        torch.randn, torch.rand, torch.normal = randn, rand, normal
        return self

    def __exit__(self, *exc):
        torch.randn, torch.rand, torch.normal = self._randn, self._rand, self._normal


def run_sampler(name, kind, arch, dtype, n_iters, n_burnin, prior_scale, seed, **kw):
    torch.manual_seed(seed)
    a = ARCHS[arch]
    data = load_data(a["data"], dtype)
    loader = DataLoader(data, batch_size=len(data))
    model = make_model(arch, dtype, prior_scale=prior_scale)
    theta0 = model.prior.sample()
    keys = ["sample", "target_val", "accepted"] + ([] if kind == "mh" else ["grad_val"])
    chain = ChainList(keys=keys)
    with Tape() as tape:
        if kind == "mala":
            s = MALA(model, theta0=theta0, dataloader=loader, step=kw["step"], chain=chain)
        elif kind == "hmc":
            tuner = HMCDATuner(l=kw["tuner_l"], e0=kw["step"], eub=kw.get("tuner_eub")) if "tuner_l" in kw else None
            s = HMC(model, theta0=theta0, dataloader=loader, step=kw["step"], num_steps=kw.get("num_steps", 1), tuner=tuner,
                    chain=chain)
        else:
            s = MetropolisHastings(model, theta0=theta0, dataloader=loader, symmetric=kw.get("symmetric", True),
                                   chain=chain)
            if "prop_scale" in kw:
                s.kernel.set_density_params(theta0.clone().detach(),
                                            scale=torch.full_like(theta0, kw["prop_scale"]))
        s.run(num_epochs=n_iters, num_burnin_epochs=n_burnin)
    z = np.stack(tape.z); u = np.concatenate(tape.u)
    assert z.shape[0] == n_iters and u.shape[0] == n_iters, (z.shape, u.shape)
    out = dict(theta0=npy(theta0), z=z, u=u, n_iters=n_iters, n_burnin=n_burnin, prior_scale=prior_scale,
               samples=npy(chain.get_samples()), target_vals=npy(chain.get_target_vals()),
               accepted=np.array(chain.vals["accepted"], dtype=np.uint8),
               final_sample=npy(s.current["sample"]), final_target=npy(s.current["target_val"]))
    if kind != "mh":
        out["grad_vals"] = npy(chain.get_grad_vals())
    for k, v in kw.items():
        out[k] = v
    if kw.get("ess"):
        out["multi_ess"] = chain.multi_ess()
    if "tuner_l" in kw:
        out["final_step"], out["final_num_steps"] = s.step, s.num_steps
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, "acceptance", out["accepted"].mean())


def adaptive_goldens():
    """The reference's AM (am.py) and RAM (ram.py) on MLP 2-2-1 and 2-3-2-1 / XOR, fp64.  AM draws a second uniform (the
    mixture test, am.py:70) only when idx + 1 > t0: the recorded u is [T, 2] with 0.5 where none was consumed."""
    from eeyore.samplers import AM, RAM
    for name, kind, arch, n_iters, n_burnin, seed, kw in (
            # without a transform the empirical covariance is singular until more than P distinct states have been visited
            # (the reference's own examples pass transform=softabs): t0 is chosen large enough for the plain estimator
            ("am_xor221_f64", "am", "221", 400, 30, 31, dict(l=0.05, b=0.6, c=0.3, t0=80)),
            ("am_xor2321_f64", "am", "2321", 500, 0, 32, dict(l=0.2, b=0.4, c=0.15, t0=200)),
            ("ram_xor221_f64", "ram", "221", 300, 30, 33, dict(a=0.234, g=0.7)),
            ("ram_xor2321_f64", "ram", "2321", 200, 0, 34, dict(a=0.3, g=0.6))):
        torch.manual_seed(seed)
        a = ARCHS[arch]
        data = load_data(a["data"], torch.float64)
        loader = DataLoader(data, batch_size=len(data))
        model = make_model(arch, torch.float64, prior_scale=math.sqrt(3.0))
        theta0 = model.prior.sample() * 0.5
        chain = ChainList(keys=["sample", "target_val", "accepted"])
        with Tape() as tape:
            s = (AM if kind == "am" else RAM)(model, theta0=theta0, dataloader=loader, chain=chain, **kw)
            s.run(num_epochs=n_iters, num_burnin_epochs=n_burnin)
        z = np.stack(tape.z)
        u_all = np.concatenate(tape.u)
        if kind == "am":
            u = np.full((n_iters, 2), 0.5)
            pos = 0
            for idx in range(n_iters):
                if idx + 1 > kw["t0"]:
                    u[idx, 0] = u_all[pos]; pos += 1
                u[idx, 1] = u_all[pos]; pos += 1
            assert pos == len(u_all)
        else:
            u = u_all
            assert len(u) == n_iters
        out = dict(theta0=npy(theta0), z=z, u=u, n_iters=n_iters, n_burnin=n_burnin,
                   samples=npy(chain.get_samples()), target_vals=npy(chain.get_target_vals()),
                   accepted=np.array(chain.vals["accepted"], dtype=np.uint8), final_sample=npy(s.current["sample"]),
                   final_factor=npy(s.cov if kind == "am" else s.chol_cov))
        out.update(kw)
        np.savez_compressed(OUT / f"{name}.npz", **out)
        print(name, "acceptance", out["accepted"].mean())


def power_posterior_goldens():
    """The reference's PowerPosteriorSampler (tempered MH / MALA chains with neighbour swaps) on MLP 2-2-1 / XOR, fp64.
    Recorded besides the proposal noise: every categorical neighbour draw (sample_categorical is wrapped on the instance)."""
    from eeyore.samplers import PowerPosteriorSampler
    for name, arch, spec_samplers, bs, n_iters, n_burnin, seed in (
            ("pp_221_mix", "221", [["MetropolisHastings", {}], ["MALA", {"step": 0.3}], ["MetropolisHastings", {}],
                                   ["MALA", {"step": 0.15}]], 3, 80, 10, 21),
            ("pp_2321_mala", "2321", [["MALA", {"step": 0.4}], ["MALA", {"step": 0.3}], ["MALA", {"step": 0.2}],
                                      ["MALA", {"step": 0.15}], ["MALA", {"step": 0.1}]], 1, 60, 0, 22)):
        torch.manual_seed(seed)
        a = ARCHS[arch]
        data = load_data(a["data"], torch.float64)
        loader = DataLoader(data, batch_size=len(data))
        model = make_model(arch, torch.float64, prior_scale=math.sqrt(3.0))
        theta0 = model.prior.sample() * 0.5
        k = len(spec_samplers)
        keys = ["sample", "target_val"]
        with Tape() as tape:
            s = PowerPosteriorSampler(model, loader, spec_samplers, theta0=theta0, between_step=bs, keys=keys)
            js = []
            orig = s.sample_categorical

            def rec(i, _orig=orig, _js=js):
                j = _orig(i)
                _js.append(j)
                return j

            s.sample_categorical = rec
            s.run(num_epochs=n_iters, num_burnin_epochs=n_burnin)
        # call order per iteration: K x (z, u) within moves; then, when idx % bs == 0, K x (categorical, u)
        nb = len(js) // k
        z = np.stack(tape.z).reshape(n_iters, k, -1)
        u_all = np.concatenate(tape.u)
        u_w = np.zeros((n_iters, k)); u_b = np.zeros((nb, k))
        pos = b_idx = 0
        for t in range(n_iters):
            u_w[t] = u_all[pos:pos + k]; pos += k
            if t % bs == 0:
                u_b[b_idx] = u_all[pos:pos + k]; pos += k; b_idx += 1
        assert pos == len(u_all) and b_idx == nb, (pos, len(u_all), b_idx, nb)
        out = dict(theta0=npy(theta0), z=z, u=u_w, j_tape=np.array(js, dtype=np.int64).reshape(nb, k), u_between=u_b,
                   n_iters=n_iters, n_burnin=n_burnin, between_step=bs, b=0.5, temperatures=np.array(s.temperature),
                   kinds=np.array([n for n, _ in spec_samplers]),
                   steps=np.array([kw.get("step", 0.0) for _, kw in spec_samplers]),
                   samples=np.stack([npy(s.samplers[i].chain.get_samples()) for i in range(k)]),
                   target_vals=np.stack([npy(s.samplers[i].chain.get_target_vals()) for i in range(k)]),
                   final_sample=np.stack([npy(s.samplers[i].current["sample"]) for i in range(k)]),
                   final_target=np.array([float(s.samplers[i].current["target_val"]) for i in range(k)]))
        np.savez_compressed(OUT / f"{name}.npz", **out)
        print(name, "between sweeps", nb, "distinct final states", len({tuple(np.round(v, 6)) for v in out["final_sample"]}))


def stats_goldens():
    out = {}
    chains = []
    for i in range(1, 5):
        c = np.loadtxt(REF / "examples" / "stats" / f"chain0{i}.csv", delimiter=",", skiprows=0)
        chains.append(c)
    x = np.stack(chains)                              # [4,1000,3]
    out["chains"] = x
    covs, inses, esss = [], [], []
    for i in range(4):
        t = torch.from_numpy(x[i])
        covs.append(npy(st.cov(t))); inses.append(npy(st.inse_mc_cov(t))); esss.append(st.multi_ess(t))
    out["cov"] = np.stack(covs); out["inse"] = np.stack(inses); out["multi_ess"] = np.array(esss)
    rhat, *_ = st.multi_rhat(torch.from_numpy(x))
    out["multi_rhat"] = np.array(complex(rhat).real if not isinstance(rhat, float) else rhat)
    np.savez_compressed(OUT / "stats_goldens.npz", **out)
    print("stats multi_ess", esss, "rhat", out["multi_rhat"])


def acf_goldens():
    """ACF lives in the un-vendored `kanga` package, so the reference cannot pin it.  These
    goldens come from two THIRD-PARTY implementations of the sample autocorrelation function over the reference's own chain
    fixtures (examples/stats/chain01..04.csv): scipy.signal.correlate (FFT method) and numpy.correlate (direct), both with
    the biased 1/n normalisation rho_k = c_k / c_0 (the convention of R's acf and of statsmodels).  Neither
    calls any code of this repository; the two must agree with each other before the file is written."""
    from scipy import signal
    x = np.stack([np.loadtxt(REF / "examples" / "stats" / f"chain0{i}.csv", delimiter=",") for i in range(1, 5)])   # [4,1000,3]
    C, n, P = x.shape
    max_lag = 40
    a_fft = np.empty((C, max_lag + 1, P))
    a_dir = np.empty_like(a_fft)
    for c in range(C):
        for j in range(P):
            v = x[c, :, j] - x[c, :, j].mean()
            full = signal.correlate(v, v, mode="full", method="fft")[n - 1:]
            a_fft[c, :, j] = full[: max_lag + 1] / full[0]
            d = np.correlate(v, v, mode="full")[n - 1:]
            a_dir[c, :, j] = d[: max_lag + 1] / d[0]
    assert np.max(np.abs(a_fft - a_dir)) < 1e-12
    np.savez_compressed(OUT / "acf_goldens.npz", acf=a_dir, acf_fft=a_fft, max_lag=np.array(max_lag))
    print("acf goldens: lag-1 of chain01", a_dir[0, 1], "max |fft - direct|", np.max(np.abs(a_fft - a_dir)))


def datapar_goldens():
    """BASELINE config 5 architecture (16-64-64-1, fp32) on a 4,099-row synthetic binary data set (ragged on purpose:
    not a multiple of the 128-row tile nor of 4): log_target and gradient from the unmodified reference."""
    rng = np.random.default_rng(4)
    n = 4099
    x = rng.normal(size=(n, 16)).astype(np.float32)
    teacher = rng.normal(size=16).astype(np.float32)
    y = ((x @ teacher + 0.5 * rng.normal(size=n)) > 0).astype(np.float32)[:, None]
    hp = Hyperparameters(dims=[16, 64, 64, 1], bias=3 * [True], activations=3 * [torch.sigmoid])
    model = MLP(loss=loss_functions["binary_classification"], hparams=hp, dtype=torch.float32)
    P = model.num_params()
    model.prior = Normal(torch.zeros(P), math.sqrt(3.0) * torch.ones(P))
    thetas = (rng.normal(size=(3, P)) * np.array([0.1, 0.3, 0.6])[:, None]).astype(np.float32)
    lts, grads = [], []
    for th in thetas:
        lt, g = model.upto_grad_log_target(torch.from_numpy(th).clone(), torch.from_numpy(x), torch.from_numpy(y))
        lts.append(lt.item()); grads.append(npy(g))
    # fp64 run of the same reference code for an error yardstick
    model64 = MLP(loss=loss_functions["binary_classification"], hparams=hp, dtype=torch.float64)
    model64.prior = Normal(torch.zeros(P, dtype=torch.float64), math.sqrt(3.0) * torch.ones(P, dtype=torch.float64))
    lts64, grads64 = [], []
    for th in thetas:
        lt, g = model64.upto_grad_log_target(torch.from_numpy(th).double(), torch.from_numpy(x).double(),
                                             torch.from_numpy(y).double())
        lts64.append(lt.item()); grads64.append(npy(g))
    np.savez_compressed(OUT / "dp_goldens.npz", x=x, y=y, theta=thetas, lt=np.array(lts), grad=np.stack(grads),
                        lt64=np.array(lts64), grad64=np.stack(grads64).astype(np.float32))
    print("dp_goldens", lts, "fp32-vs-fp64 grad rel err",
          [float(np.abs(grads[i] - grads64[i]).max() / np.abs(grads64[i]).max()) for i in range(3)])


def model_goldens_f32_inputs_in_f64():
    """The fp32 fixtures of model_goldens.npz (theta and data rounded to fp32) evaluated by the unmodified reference in
    fp64: the yardstick BASELINE.json's 1e-5 fp32 bar is measured against (an fp32 evaluation by torch carries its own
    ~1e-6 summation noise, which is not the kernel's error)."""
    mg = np.load(OUT / "model_goldens.npz")
    out = {}
    for arch, a in ARCHS.items():
        data = load_data(a["data"], torch.float32)
        x64, y64 = data.x.double(), data.y.double()
        for ps, pst in ((1.0, "p1"), (100.0, "p100"), (math.sqrt(3.0), "psqrt3")):
            for temp, tt in ((None, ""), (0.7, "_T07")):
                key = f"{arch}_f32_{pst}{tt}"
                model = make_model(arch, torch.float64, prior_scale=float(np.float32(ps)), temperature=temp)
                lts, grads, lls, lps = [], [], [], []
                for th in torch.from_numpy(mg[key + "_theta"]).double():
                    lt, g = model.upto_grad_log_target(th.clone().detach(), x64, y64)
                    lts.append(lt.item()); grads.append(npy(g))
                    lls.append(model.log_lik(x64, y64).item()); lps.append(model.log_prior().item())
                out[key + "_lt64"] = np.array(lts); out[key + "_grad64"] = np.stack(grads)
                out[key + "_ll64"] = np.array(lls); out[key + "_lp64"] = np.array(lps)
    np.savez_compressed(OUT / "model_goldens_f32ref.npz", **out)
    print("model_goldens_f32ref", len(out))


def smmala_goldens():
    """Independent pin of the (builder-defined) SMMALA: the snapshot has no SMMALA sampler, but it ships
    every piece one is made of.  This run uses ONLY those pieces and torch -- none of the builder's restatement:
      * log-target and gradient: the reference MLP, upto_grad_log_target (models/log_target_model.py:15-23);
      * metric: expected Fisher information from per-row autograd derivatives of the reference MLP's output
        (d p_i / d theta through MLP.forward, mlp.py:45-50): sum_i (dp_i)(dp_i)^T / (p_i (1 - p_i)) + prior precision;
      * positive-definiteness test: eeyore.linalg.is_pos_def (linalg/is_pos_def.py:3-11); factor: torch.linalg.cholesky;
      * proposal density: eeyore.kernels.MultivariateNormalKernel(loc, scale_tril).log_prob
        (kernels/multivariate_normal_kernel.py:5-19, normalized_kernel.py:14-16), covariance step * G^-1;
      * accept test: the MALA structure of samplers/mala.py:58-66, log(u) < log_rate.
    The draw theta' = mean + sqrt(step) R^-T z (G = R R^T) is A.7's convention for mapping the recorded z to a proposal."""
    from eeyore.kernels import MultivariateNormalKernel
    from eeyore.linalg import is_pos_def
    dt = torch.float64
    s3 = math.sqrt(3.0)
    rng = np.random.default_rng(3)
    corners = np.array([[0, 0], [0, 1], [1, 0], [1, 1]], dtype=np.float64)
    nx = np.concatenate([c + 0.15 * rng.normal(size=(10, 2)) for c in corners])
    ny = np.concatenate([np.full((10, 1), float(int(c[0]) ^ int(c[1]))) for c in corners])
    xor = load_data("xor", dt)
    cases = (("smmala_xor2321_f64", "2321", npy(xor.x), npy(xor.y), 0.9, 70, 10, 41),
             ("smmala_nxor2321_f64", "2321", nx, ny, 0.12, 60, 0, 42),
             ("smmala_xor221_f64", "221", npy(xor.x), npy(xor.y), 1.3, 90, 20, 43))
    for name, arch, x_np, y_np, step, n_iters, n_burnin, seed in cases:
        torch.manual_seed(seed)
        model = make_model(arch, dt, prior_scale=s3)
        P = model.num_params()
        x, y = torch.from_numpy(x_np), torch.from_numpy(y_np)
        prec = 1.0 / model.prior.scale ** 2

        def state_at(theta):
            lt, g = model.upto_grad_log_target(theta.clone().detach(), x, y)
            p = model(x)[:, 0]
            G = torch.diag(prec.clone())
            for i in range(x.shape[0]):
                dp = torch.cat([v.reshape(-1) for v in torch.autograd.grad(p[i], list(model.parameters()), retain_graph=True)])
                G = G + torch.outer(dp, dp) / (p[i] * (1 - p[i])).detach()
            G = (G + G.t()) / 2
            lt, g, G = lt.detach(), g.detach(), G.detach()
            if not (torch.isfinite(G).all() and is_pos_def(G)):
                return None
            R = torch.linalg.cholesky(G)
            mean = theta + 0.5 * step * torch.linalg.solve(G, g)
            cov = step * torch.linalg.inv(G)
            kern = MultivariateNormalKernel(mean, torch.linalg.cholesky((cov + cov.t()) / 2))
            return dict(theta=theta, lt=lt, g=g, R=R, mean=mean, kern=kern)

        theta0 = model.prior.sample() * 0.5
        cur = state_at(theta0)
        assert cur is not None
        zs, us, samples, lts, grads, accs = [], [], [], [], [], []
        for t in range(n_iters):
            z = torch.randn(P, dtype=dt)
            u = torch.rand(1, dtype=dt)
            prop_theta = cur["mean"] + math.sqrt(step) * torch.linalg.solve_triangular(cur["R"].t(), z[:, None], upper=True)[:, 0]
            prop = state_at(prop_theta)
            acc = False
            if prop is not None:
                log_rate = prop["lt"] - cur["lt"] - cur["kern"].log_prob(prop_theta) + prop["kern"].log_prob(cur["theta"])
                acc = bool(torch.log(u) < log_rate)
            if acc:
                cur = prop
            zs.append(npy(z)); us.append(u.item())
            if t >= n_burnin:
                samples.append(npy(cur["theta"])); lts.append(cur["lt"].item()); grads.append(npy(cur["g"])); accs.append(int(acc))
        out = dict(x=x_np, y=y_np, theta0=npy(theta0), z=np.stack(zs), u=np.array(us), step=step, n_iters=n_iters,
                   n_burnin=n_burnin, prior_scale=s3, samples=np.stack(samples), target_vals=np.array(lts),
                   grad_vals=np.stack(grads), accepted=np.array(accs, dtype=np.uint8), final_sample=npy(cur["theta"]),
                   final_target=cur["lt"].item())
        np.savez_compressed(OUT / f"{name}.npz", **out)
        print(name, "acceptance", out["accepted"].mean())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "smmala":
        smmala_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "acf":
        acf_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "f32ref":
        model_goldens_f32_inputs_in_f64()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dp":
        datapar_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "adaptive":
        adaptive_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pp":
        power_posterior_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "tuner":
        s3 = math.sqrt(3.0)
        run_sampler("hmcda_xor2321_f64", "hmc", "2321", torch.float64, 90, 50, s3, 7, step=0.05, tuner_l=0.6)
        run_sampler("hmcda_iris433_f64", "hmc", "433", torch.float64, 70, 40, s3, 8, step=0.01, tuner_l=0.15, tuner_eub=0.05)
        sys.exit(0)
    model_goldens()
    datapar_goldens()
    s3 = math.sqrt(3.0)
    # BASELINE.json configs[0]: MLP 2-2-1 XOR, single MALA chain, 1100 iterations (110 burn-in), fp64
    run_sampler("mala_xor221_f64", "mala", "221", torch.float64, 1100, 110, s3, 0, step=1.74, ess=True)
    run_sampler("mala_iris433_f32", "mala", "433", torch.float32, 120, 20, s3, 1, step=0.003)
    run_sampler("mala_iris433_f64", "mala", "433", torch.float64, 120, 20, s3, 1, step=0.003)
    run_sampler("hmc_xor2321_f64", "hmc", "2321", torch.float64, 160, 20, s3, 2, step=0.3, num_steps=10)
    run_sampler("hmc_xor221_f64", "hmc", "221", torch.float64, 160, 20, s3, 3, step=0.9, num_steps=7)
    run_sampler("hmc_xor2321_f64_s09", "hmc", "2321", torch.float64, 160, 20, s3, 2, step=0.9, num_steps=10)
    run_sampler("hmc_iris433_f64", "hmc", "433", torch.float64, 60, 10, s3, 4, step=0.04, num_steps=10)
    run_sampler("hmc_iris433_f32", "hmc", "433", torch.float32, 40, 10, s3, 4, step=0.02, num_steps=10)
    run_sampler("hmcda_xor2321_f64", "hmc", "2321", torch.float64, 90, 50, s3, 7, step=0.05, tuner_l=0.6)
    run_sampler("hmcda_iris433_f64", "hmc", "433", torch.float64, 70, 40, s3, 8, step=0.01, tuner_l=0.15, tuner_eub=0.05)
    run_sampler("mh_xor221_f64", "mh", "221", torch.float64, 400, 50, s3, 5)
    run_sampler("mh_xor2321_f64_nonsym", "mh", "2321", torch.float64, 300, 0, s3, 6, symmetric=False,
                prop_scale=0.4)
    stats_goldens()
    acf_goldens()
    power_posterior_goldens()
    adaptive_goldens()
    model_goldens_f32_inputs_in_f64()
    smmala_goldens()
