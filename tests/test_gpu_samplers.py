"""GPU parity: fused MH / MALA / HMC kernels.  Fed the reference's proposal noise (golden tapes recorded from the
unmodified reference by oracle/make_golden.py) they must reproduce its accept decisions and states."""
import numpy as np
import pytest
import torch

import oracle
from eeyore_b200.chains import ChainList, DeviceChains
from eeyore_b200.samplers import HMC, MALA, MetropolisHastings
from gpu_helpers import dataset, loader, make_model, npy, T_DTYPES
from helpers import NP_DTYPES, data_of, load, rel_err, spec_of

pytestmark = pytest.mark.gpu

GOLDEN_RUNS = [
    ("mala_xor221_f64", MALA, dict(step=1.74)),
    ("mala_iris433_f64", MALA, dict(step=0.003)),
    ("hmc_xor2321_f64", HMC, dict(step=0.3, num_steps=10)),
    ("hmc_xor2321_f64_s09", HMC, dict(step=0.9, num_steps=10)),
    ("hmc_xor221_f64", HMC, dict(step=0.9, num_steps=7)),
    ("hmc_iris433_f64", HMC, dict(step=0.04, num_steps=10)),
    ("mh_xor221_f64", MetropolisHastings, dict()),
    ("mh_xor2321_f64_nonsym", MetropolisHastings, dict(symmetric=False, scale=0.4)),
]


def _arch(name):
    return name.split("_")[1].replace("xor", "").replace("iris", "")


@pytest.mark.parametrize("name,cls,kw", GOLDEN_RUNS)
@pytest.mark.parametrize("lanes", [0, 1, 4, 32])
def test_reference_trajectories(name, cls, kw, lanes):
    """sampler.run(num_epochs, num_burnin_epochs) through the reference-shaped API; one fused launch."""
    gd = load(name)
    arch = _arch(name)
    m = make_model(arch, "f64", float(gd["prior_scale"]))
    ds = dataset(arch, "f64")
    keys = ["sample", "target_val", "accepted"] + ([] if cls is MetropolisHastings else ["grad_val"])
    s = cls(m, theta0=torch.from_numpy(gd["theta0"]), dataloader=loader(ds), chain=ChainList(keys=keys),
            lanes_per_chain=lanes, **kw)
    s.set_noise_tape(torch.from_numpy(gd["z"]), torch.from_numpy(gd["u"]))
    s.run(num_epochs=int(gd["n_iters"]), num_burnin_epochs=int(gd["n_burnin"]))
    ch = s.get_chain()
    assert len(ch) == int(gd["n_iters"]) - int(gd["n_burnin"])
    assert np.array_equal(np.array(ch.vals["accepted"], dtype=np.uint8), gd["accepted"]), "accept decisions differ"
    assert rel_err(npy(ch.get_samples()), gd["samples"]) < 1e-10
    assert rel_err(npy(ch.get_target_vals()), gd["target_vals"]) < 1e-10
    if "grad_val" in keys:
        assert rel_err(npy(ch.get_grad_vals()), gd["grad_vals"]) < 1e-9
    assert rel_err(npy(s.current["sample"]), gd["final_sample"]) < 1e-10
    assert abs(s.current["target_val"].item() - float(gd["final_target"])) < 1e-9 * abs(float(gd["final_target"]))
    assert s.current["accepted"] == int(gd["accepted"][-1])
    assert s.counter.idx == int(gd["n_iters"])
    assert abs(ch.acceptance_rate() - gd["accepted"].mean()) < 1e-12


def test_config1_mala_draw_by_draw():
    """BASELINE configs[0] through the per-iteration entry point (draw), as SerialSampler.run drives it."""
    gd = load("mala_xor221_f64")
    m = make_model("221", "f64", float(gd["prior_scale"]))
    ds = dataset("221", "f64")
    s = MALA(m, theta0=torch.from_numpy(gd["theta0"]), dataloader=loader(ds), step=1.74)
    s.set_noise_tape(torch.from_numpy(gd["z"]), torch.from_numpy(gd["u"]))
    nb, n = int(gd["n_burnin"]), 200
    acc = []
    for t in range(n):
        s.draw(ds.x, ds.y, savestate=t >= nb)
        acc.append(s.current["accepted"])
    assert acc[nb:] == gd["accepted"][: n - nb].tolist()
    assert rel_err(npy(s.get_chain().get_samples()), gd["samples"][: n - nb]) < 1e-10


@pytest.mark.parametrize("name,cls,kw,steps", [("mala_iris433_f32", MALA, dict(step=0.003), 10),
                                               ("hmc_iris433_f32", HMC, dict(step=0.02, num_steps=10), 4)])
def test_fp32_trajectory_prefix(name, cls, kw, steps):
    """fp32: rounding differences between summation orders grow along a chain, so compare a short prefix at 1e-3 and
    single evaluations at 1e-5 (test_gpu_log_target)."""
    gd = load(name)
    m = make_model("433", "f32", float(gd["prior_scale"]))
    ds = dataset("433", "f32")
    s = cls(m, theta0=torch.from_numpy(gd["theta0"]), dataloader=loader(ds), **kw)
    s.set_noise_tape(torch.from_numpy(gd["z"]), torch.from_numpy(gd["u"]))
    nb = int(gd["n_burnin"])
    s.run(num_epochs=nb + steps, num_burnin_epochs=nb)
    assert rel_err(npy(s.get_chain().get_samples()), gd["samples"][:steps]) < 1e-3


@pytest.mark.parametrize("kind,arch,kw", [("hmc", "2321", dict(step=0.5, num_steps=5)),
                                          ("hmc", "433", dict(step=0.03, num_steps=4)),
                                          ("mala", "4323", dict(step=0.002)),
                                          ("mala", "221", dict(step=1.2)),
                                          ("mh", "2321", dict(scale=0.3))])
@pytest.mark.parametrize("lanes", [1, 8, 16])
def test_many_chains_vs_oracle(kind, arch, kw, lanes):
    """C chains side by side with an arbitrary noise tape, against the chain-batched oracle (ragged C: not a multiple
    of the block's chains)."""
    dt = np.float64
    rng = np.random.default_rng(11)
    C, T, nb = 45, 14, 3
    spec = spec_of(arch)
    P = spec.num_params
    x, y = data_of(arch, dt)
    s3 = 3 ** 0.5
    theta0 = rng.normal(size=(C, P)) * (0.5 if arch in ("433", "4323") else 1.5)
    z, u = rng.normal(size=(T, C, P)), rng.uniform(size=(T, C))
    m = make_model(arch, "f64", s3)
    ds = dataset(arch, "f64")
    cls = dict(hmc=HMC, mala=MALA, mh=MetropolisHastings)[kind]
    s = cls(m, theta0=torch.from_numpy(theta0), dataloader=loader(ds), lanes_per_chain=lanes, **kw)
    s.set_noise_tape(torch.from_numpy(z), torch.from_numpy(u))
    s.run(num_epochs=T, num_burnin_epochs=nb)
    got = s.get_chain()
    assert isinstance(got, DeviceChains) and got.num_chains() == C and got.num_samples() == T - nb
    loc, scale = np.zeros(P), np.full(P, s3)
    if kind == "hmc":
        ref = oracle.hmc_run(spec, x, y, loc, scale, theta0, z, u, kw["step"], kw["num_steps"], n_burnin=nb)
    elif kind == "mala":
        ref = oracle.mala_run(spec, x, y, loc, scale, theta0, z, u, kw["step"], n_burnin=nb)
    else:
        ref = oracle.mh_run(spec, x, y, loc, scale, theta0, z, u, n_burnin=nb, prop_scale=kw["scale"])
    assert np.array_equal(npy(got.accepted_soa), ref["accepted"])
    assert rel_err(npy(got.get_samples().permute(1, 0, 2)), ref["sample"]) < 1e-10
    assert rel_err(npy(got.target_soa), ref["target_val"]) < 1e-10
    assert np.array_equal(npy(s.acceptance_counts()) >= ref["accepted"].sum(0), np.ones(C, bool))
    assert rel_err(npy(s.current["sample"]), ref["final"]["sample"]) < 1e-10


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_philox_draws_match_the_pinned_stream(tag):
    import ctypes as C
    from eeyore_b200 import _native as nv
    dt = T_DTYPES[tag]
    Cn, P, seed, it, c0 = 1000, 27, 0xABCDEF0123456789, 77, 5
    z = torch.empty(Cn, P, dtype=dt, device="cuda")
    u = torch.empty(Cn, dtype=dt, device="cuda")
    nv.check(nv.lib().eeyore_b200_philox_draws(nv.DTYPE_IDS[dt], Cn, P, seed, it, c0, nv.ptr(z), nv.ptr(u), None))
    torch.cuda.synchronize()
    zr = oracle.chain_normals(seed, np.arange(c0, c0 + Cn), it, P, NP_DTYPES[tag])
    ur = oracle.chain_uniforms(seed, np.arange(c0, c0 + Cn), it, NP_DTYPES[tag])
    assert np.array_equal(npy(u), ur)
    assert np.max(np.abs(npy(z) - zr)) < (1e-11 if tag == "f64" else 1e-4)


def test_philox_mode_hmc_matches_oracle_and_is_shard_invariant():
    """Production RNG: on-device Philox keyed by (seed, global chain id, iteration).  The oracle gets the CPU replica of
    the stream; running a slice of the chains with chain_offset reproduces the same chains (chain sharding)."""
    arch, P, C, T, L, step, seed = "2321", 20, 70, 9, 5, 0.4, 2024
    s3 = 3 ** 0.5
    rng = np.random.default_rng(3)
    theta0 = rng.normal(size=(C, P))
    m = make_model(arch, "f64", s3)
    ds = dataset(arch, "f64")
    s = HMC(m, theta0=torch.from_numpy(theta0), dataloader=loader(ds), step=step, num_steps=L, seed=seed)
    s.run(num_epochs=T, num_burnin_epochs=0)
    got = s.get_chain()
    z = np.stack([oracle.chain_normals(seed, np.arange(C), t, P) for t in range(T)])
    u = np.stack([oracle.chain_uniforms(seed, np.arange(C), t) for t in range(T)])
    x, y = data_of(arch, np.float64)
    ref = oracle.hmc_run(spec_of(arch), x, y, np.zeros(P), np.full(P, s3), theta0, z, u, step, L)
    assert np.array_equal(npy(got.accepted_soa), ref["accepted"])
    assert rel_err(npy(got.get_samples().permute(1, 0, 2)), ref["sample"]) < 1e-9
    # shard: chains 40..69 alone, told their global ids
    s2 = HMC(m, theta0=torch.from_numpy(theta0[40:]), dataloader=loader(ds), step=step, num_steps=L, seed=seed)
    s2.chain_offset = 40
    s2.run(num_epochs=T, num_burnin_epochs=0)
    assert torch.equal(s2.get_chain().samples_soa, got.samples_soa[:, :, 40:])
    # as in the reference's loop, run() performs num_epochs MORE draws and the counter keeps counting: 4 + 5 iterations ==
    # 9 iterations (same Philox stream)
    s3_ = HMC(m, theta0=torch.from_numpy(theta0), dataloader=loader(ds), step=step, num_steps=L, seed=seed)
    s3_.run(num_epochs=4, num_burnin_epochs=0)
    s3_.run(num_epochs=5, num_burnin_epochs=0)
    assert s3_.counter.idx == 9
    assert torch.equal(s3_.get_chain().samples_soa, got.samples_soa)


@pytest.mark.parametrize("layout", ["npc", "cnp"])
def test_host_output_saved_states_equal_the_device_blocks(layout):
    """host_output: the kernel stores the saved states straight into pinned host memory (no staging copy).  Same seed => the
    host buffers hold bit for bit what the device blocks of an ordinary run hold, across two runs re-using the buffers."""
    arch, P, C, T, seed = "2321", 20, 333, 12, 77
    theta0 = torch.from_numpy(np.random.default_rng(5).normal(size=(C, P)))
    m = make_model(arch, "f64", 3 ** 0.5)
    ds = dataset(arch, "f64")
    runs = {}
    for host in (False, True):
        s = HMC(m, theta0=theta0, dataloader=loader(ds), step=0.3, num_steps=4, seed=seed, thin=3)
        s.sample_layout = layout
        s.host_output = host
        blocks = []
        for rep in range(2):
            s.reset(theta0 + rep)
            s.run(num_epochs=T, num_burnin_epochs=2)
            torch.cuda.synchronize()
            b = s._device_blocks[-1]
            assert b["sample"].device.type == ("cpu" if host else "cuda")
            assert not host or b["sample"].is_pinned() or b["sample"].permute(2, 0, 1).is_pinned()
            blocks.append({k: v.cpu().clone() for k, v in b.items()})
        runs[host] = blocks
    for rep in range(2):
        for k in ("sample", "target_val", "accepted"):
            assert torch.equal(runs[True][rep][k], runs[False][rep][k]), (rep, k)
    assert not torch.equal(runs[True][0]["sample"], runs[True][1]["sample"])
    # the final state of the last run, written by the kernel into pinned host memory beside the device state
    fin = s.host_current
    assert fin is not None and fin["sample"].device.type == "cpu"
    assert torch.equal(fin["sample"], s.current["sample"].cpu())
    assert torch.equal(fin["target_val"], s.current["target_val"].cpu())
    assert torch.equal(fin["accept_count"], s.acceptance_counts().cpu().to(torch.int32))


def test_nan_proposals_are_rejected():
    """a saturated proposal gives a NaN target; comparisons with NaN are False => reject, state intact."""
    m = make_model("221", "f32", 1.0)
    ds = dataset("221", "f32")
    theta0 = torch.tensor([[0.1, -0.2, 0.3, 0.1, 0.0, 0.2, 0.0, 0.0, 3.0]] * 4, dtype=torch.float32)
    T = 6
    z = torch.zeros(T, 4, 9, dtype=torch.float32)
    z[:, :, 8] = 100.0          # pushes the output bias far into saturation: p == 1.0f exactly
    u = torch.full((T, 4), 0.5, dtype=torch.float32)
    for cls, kw in ((MetropolisHastings, dict()), (MALA, dict(step=0.5)), (HMC, dict(step=0.5, num_steps=3))):
        s = cls(m, theta0=theta0, dataloader=loader(ds), **kw)
        s.set_noise_tape(z, u)
        s.run(num_epochs=T, num_burnin_epochs=0)
        got = s.get_chain()
        assert got.accepted_soa.sum().item() == 0, cls.__name__
        assert torch.equal(got.get_samples()[:, -1].cpu(), theta0)
        assert torch.isfinite(s.current["target_val"]).all()


def test_thinning_and_sample_layout():
    m = make_model("2321", "f64", 1.0)
    ds = dataset("2321", "f64")
    theta0 = torch.randn(33, 20, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    a = HMC(m, theta0=theta0, dataloader=loader(ds), step=0.3, num_steps=3, seed=5)
    a.run(num_epochs=20, num_burnin_epochs=5)
    b = HMC(m, theta0=theta0, dataloader=loader(ds), step=0.3, num_steps=3, seed=5, thin=4)
    b.run(num_epochs=20, num_burnin_epochs=5)
    full, thin = a.get_chain(), b.get_chain()
    assert full.num_samples() == 15 and thin.num_samples() == 4
    assert torch.equal(thin.samples_soa, full.samples_soa[0::4])
    assert torch.equal(thin.get_samples(), full.get_samples()[:, 0::4])
    cl = full.to_chainlist(7)
    assert torch.equal(cl.get_samples(), full.get_samples()[7]) and len(cl.vals["accepted"]) == 15


def test_full_size_hmc_config4_properties():
    """BASELINE config 4 size: 524,288 chains of HMC on 2-3-2-1 XOR.  Properties: every chain's stored target equals a
    fresh evaluation at its stored sample; chains are reproducible; a random subset matches the oracle."""
    m = make_model("2321", "f64", 3 ** 0.5)
    ds = dataset("2321", "f64")
    C, T, L, step, seed = 524288, 3, 10, 0.3, 7
    g = torch.Generator(device="cuda").manual_seed(0)
    theta0 = torch.randn(C, 20, dtype=torch.float64, device="cuda", generator=g)
    s = HMC(m, theta0=theta0, dataloader=loader(ds), step=step, num_steps=L, seed=seed)
    s.run(num_epochs=T, num_burnin_epochs=0)
    got = s.get_chain()
    last = got.samples_soa[-1].t().contiguous()
    lt, _ = m._eval(last, *s._data_dev, want_grad=False)
    assert torch.allclose(lt, got.target_soa[-1], rtol=1e-12, atol=0)
    acc = got.acceptance().mean().item()
    assert 0.8 < acc <= 1.0
    idx = np.array([0, 1, 31, 32, 127, 128, 4095, 65535, 65536, 262143, 524287])
    z = np.stack([oracle.chain_normals(seed, idx, t, 20) for t in range(T)])
    u = np.stack([oracle.chain_uniforms(seed, idx, t) for t in range(T)])
    x, y = data_of("2321", np.float64)
    ref = oracle.hmc_run(spec_of("2321"), x, y, np.zeros(20), np.full(20, 3 ** 0.5), npy(theta0[idx]), z, u, step, L)
    assert np.array_equal(npy(got.accepted_soa[:, idx]), ref["accepted"])
    assert rel_err(npy(got.get_samples()[idx].permute(1, 0, 2)), ref["sample"]) < 1e-9


@pytest.mark.parametrize("name,arch,l,e0,eub", [("hmcda_xor2321_f64", "2321", 0.6, 0.05, None),
                                                 ("hmcda_iris433_f64", "433", 0.15, 0.01, 0.05)])
def test_hmc_with_hmcda_tuner_reference_trajectories(name, arch, l, e0, eub):
    """HMC + HMCDATuner (hmc.py:17-28,158-163; tuners/hmcda_tuner.py:8-59), adaptation on device."""
    from eeyore_b200.tuners import HMCDATuner
    gd = load(name)
    m = make_model(arch, "f64", float(gd["prior_scale"]))
    ds = dataset(arch, "f64")
    s = HMC(m, theta0=torch.from_numpy(gd["theta0"]), dataloader=loader(ds), tuner=HMCDATuner(l=l, e0=e0, eub=eub),
            chain=ChainList(keys=["sample", "target_val", "accepted", "grad_val"]))
    assert s.step == e0 and s.num_steps == max(1, round(l / e0))
    s.set_noise_tape(torch.from_numpy(gd["z"]), torch.from_numpy(gd["u"]))
    s.run(num_epochs=int(gd["n_iters"]), num_burnin_epochs=int(gd["n_burnin"]))
    ch = s.get_chain()
    assert np.array_equal(np.array(ch.vals["accepted"], dtype=np.uint8), gd["accepted"])
    assert rel_err(npy(ch.get_samples()), gd["samples"]) < 1e-9
    assert abs(s.step - float(gd["final_step"])) < 1e-11 * float(gd["final_step"])
    assert s.num_steps == int(gd["final_num_steps"])


def test_hmcda_tuner_many_chains_vs_oracle_and_split_runs():
    from eeyore_b200.tuners import HMCDATuner
    arch, P, C, T, nb, l, e0 = "2321", 20, 21, 40, 25, 0.5, 0.08
    rng = np.random.default_rng(4)
    theta0 = rng.normal(size=(C, P))
    z, u = rng.normal(size=(T, C, P)), rng.uniform(size=(T, C))
    s3 = 3 ** 0.5
    x, y = data_of(arch, np.float64)
    ref = oracle.hmc_run(spec_of(arch), x, y, np.zeros(P), np.full(P, s3), theta0, z, u, e0, 1, n_burnin=nb,
                         tuner=oracle.DATuner(l=l, e0=e0, n_chains=C))
    m = make_model(arch, "f64", s3)
    ds = dataset(arch, "f64")
    for lanes in (1, 4):
        s = HMC(m, theta0=torch.from_numpy(theta0), dataloader=loader(ds), tuner=HMCDATuner(l=l, e0=e0), lanes_per_chain=lanes)
        s.set_noise_tape(torch.from_numpy(z), torch.from_numpy(u))
        s.run(num_epochs=T, num_burnin_epochs=nb)
        got = s.get_chain()
        assert np.array_equal(npy(got.accepted_soa), ref["accepted"])
        assert rel_err(npy(got.get_samples().permute(1, 0, 2)), ref["sample"]) < 1e-8
        assert np.allclose(npy(s.step), ref["final"]["step"], rtol=1e-10)
        assert np.array_equal(npy(s.num_steps), ref["final"]["num_steps"])
    assert len(np.unique(ref["final"]["num_steps"])) > 1 or len(np.unique(np.round(ref["final"]["step"], 6))) > 1
    # Philox mode: burn-in split over two run() calls == one call (tuner iteration index and burn-in window carry over)
    a = HMC(m, theta0=torch.from_numpy(theta0), dataloader=loader(ds), tuner=HMCDATuner(l=l, e0=e0), seed=9)
    a.run(num_epochs=T, num_burnin_epochs=nb)
    b = HMC(m, theta0=torch.from_numpy(theta0), dataloader=loader(ds), tuner=HMCDATuner(l=l, e0=e0), seed=9)
    b.run(num_epochs=10, num_burnin_epochs=nb)      # 10 tuning iterations ...
    b.run(num_epochs=T - 10, num_burnin_epochs=nb)  # ... then the remaining 30 (15 still tuning)
    assert torch.equal(a.get_chain().samples_soa, b.get_chain().samples_soa)
    assert torch.equal(a.step, b.step)


def test_chainfile_backed_sampler_writes_what_a_chainlist_holds(tmp_path):
    """A sampler constructed with chain=ChainFile(...) (eeyore/chains/chain_file.py:28-45): the fused run appends its saved
    states to <key>.csv; reading the files back gives the ChainList of the same run, '%.18e' round-trip exact."""
    from eeyore_b200.chains import ChainFile
    gd = load("mala_xor221_f64")
    m = make_model("221", "f64", float(gd["prior_scale"]))
    ds = dataset("221", "f64")
    keys = ["sample", "target_val", "grad_val", "accepted"]
    runs = {}
    for kind in ("list", "file"):
        chain = ChainList(keys=keys) if kind == "list" else ChainFile(keys=keys, path=tmp_path / "mala", mode="w")
        s = MALA(m, theta0=torch.from_numpy(gd["theta0"]), dataloader=loader(ds), chain=chain, step=1.74)
        s.set_noise_tape(torch.from_numpy(gd["z"]), torch.from_numpy(gd["u"]))
        s.run(num_epochs=200, num_burnin_epochs=20)
        runs[kind] = s
    back = runs["file"].get_chain().to_chainlist(keys=keys)
    want = runs["list"].get_chain()
    assert len(back) == len(want) == 180
    assert torch.equal(back.get_samples(), want.get_samples().cpu())
    assert torch.equal(back.get_target_vals(), want.get_target_vals().cpu())
    assert torch.equal(back.get_grad_vals(), want.get_grad_vals().cpu())
    assert back.vals["accepted"] == want.vals["accepted"] and 0 < sum(back.vals["accepted"]) < 180
    # a second run() appends (mode 'a' semantics of ChainFile.reset re-opening the files)
    runs["file"].chain.mode = "a"
    runs["file"].run(num_epochs=10, num_burnin_epochs=20)
    assert len(runs["file"].get_chain().to_chainlist(keys=keys)) == 190


def test_reset_reuses_the_device_buffers_and_restarts_the_chain():
    """reset(theta_host) with the same number of chains keeps every device buffer (the e2e path of bench.py) and gives the
    same chains as a freshly constructed sampler."""
    m = make_model("2321", "f64", 3.0 ** 0.5)
    ds = dataset("2321", "f64")
    g = torch.Generator().manual_seed(3)
    th_a = torch.randn(300, m.num_params(), dtype=torch.float64, generator=g)
    th_b = torch.randn(300, m.num_params(), dtype=torch.float64, generator=g).pin_memory()
    s = HMC(m, theta0=th_a, dataloader=loader(ds), step=0.3, num_steps=5, seed=21)
    s.run(num_epochs=6, num_burnin_epochs=0)
    ptrs = (s._theta_soa.data_ptr(), s._grad_soa.data_ptr(), s._lt.data_ptr())
    s._iter_offset = 0
    s.reset(th_b)
    assert ptrs == (s._theta_soa.data_ptr(), s._grad_soa.data_ptr(), s._lt.data_ptr())
    assert s.counter.idx == 0 and int(s.acceptance_counts().sum()) == 0
    s.run(num_epochs=6, num_burnin_epochs=0)
    fresh = HMC(m, theta0=th_b, dataloader=loader(ds), step=0.3, num_steps=5, seed=21)
    fresh.run(num_epochs=6, num_burnin_epochs=0)
    assert torch.equal(s.get_chain().samples_soa, fresh.get_chain().samples_soa)
    assert torch.equal(s.acceptance_counts(), fresh.acceptance_counts())
    # another number of chains re-allocates
    s.reset(th_a[:17])
    assert s.num_chains == 17 and s._theta_soa.shape == (m.num_params(), 17)


def test_tuner_state_does_not_outlive_its_chains():
    """reset() with another number of chains after a batched tuned run: the [4, C] dual-averaging buffer is rebuilt (it
    would be indexed out of bounds otherwise) and the scalar initial step is restored; _spawn works after a batched run."""
    from eeyore_b200.tuners import HMCDATuner
    m = make_model("2321", "f64", 3.0 ** 0.5)
    ds = dataset("2321", "f64")
    th = torch.randn(96, m.num_params(), dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    s = HMC(m, theta0=th, dataloader=loader(ds), tuner=HMCDATuner(l=1.5, e0=0.25), seed=2)
    s.run(num_epochs=30, num_burnin_epochs=20)
    assert isinstance(s.step, torch.Tensor) and s.step.shape == (96,)
    child = s._spawn(th[:8])
    assert child.step == 0.25 and child.tuner.e0 == 0.25
    s.reset(th[:40])
    assert s.step == 0.25 and s._tuner_state is None
    s._iter_offset = 0                    # rewind the Philox iteration counter as well: same noise as a fresh sampler
    s.run(num_epochs=30, num_burnin_epochs=20)
    assert s.step.shape == (40,) and s._tuner_state.shape == (4, 40)
    ref = HMC(m, theta0=th[:40], dataloader=loader(ds), tuner=HMCDATuner(l=1.5, e0=0.25), seed=2)
    ref.run(num_epochs=30, num_burnin_epochs=20)
    assert torch.equal(ref.step, s.step)
