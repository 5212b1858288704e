"""GPU parity: SMMALA (Fisher metric, warp-level Cholesky).  The reference snapshot has no SMMALA sampler;
the algorithm is pinned by tests/golden/smmala_*.npz, runs assembled in oracle/make_golden.py:smmala_goldens from the
reference's own pieces only (its MLP + autograd row derivatives for the Fisher metric, is_pos_def, torch.linalg.cholesky,
MultivariateNormalKernel.log_prob, MALA's accept structure); the oracle restatement covers the sizes beyond them."""
import numpy as np
import pytest
import torch

import oracle
from eeyore_b200.datasets import XYDataset
from eeyore_b200.samplers import SMMALA
from gpu_helpers import dataset, loader, make_model, npy
from helpers import data_of, rel_err, spec_of

pytestmark = pytest.mark.gpu
S3 = 3 ** 0.5


def noisy_xor(n_per_corner=50, seed=3):
    """BASELINE config 3 data shape: 50 points per XOR corner, x = corner + N(0, 0.15^2)  (SURVEY.md 8(d))."""
    rng = np.random.default_rng(seed)
    corners = np.array([[0, 0], [0, 1], [1, 0], [1, 1]], dtype=np.float64)
    x = np.concatenate([c + 0.15 * rng.normal(size=(n_per_corner, 2)) for c in corners])
    y = np.concatenate([np.full((n_per_corner, 1), float(int(c[0]) ^ int(c[1]))) for c in corners])
    return x, y


@pytest.mark.parametrize("name,arch", [("smmala_xor2321_f64", "2321"), ("smmala_nxor2321_f64", "2321"),
                                       ("smmala_xor221_f64", "221")])
def test_smmala_reference_pieces_golden(name, arch):
    """The kernel fed the golden run's noise: identical accept decisions, states / targets / gradients < 1e-9."""
    from helpers import load
    gd = load(name)
    m = make_model(arch, "f64", float(gd["prior_scale"]))
    ds = XYDataset(torch.from_numpy(gd["x"]), torch.from_numpy(gd["y"]))
    from eeyore_b200.chains import ChainList
    s = SMMALA(m, theta0=torch.from_numpy(gd["theta0"]), dataloader=loader(ds), step=float(gd["step"]),
               chain=ChainList(keys=["sample", "target_val", "grad_val", "accepted"]))
    s.set_noise_tape(torch.from_numpy(gd["z"]), torch.from_numpy(gd["u"]))
    s.run(num_epochs=int(gd["n_iters"]), num_burnin_epochs=int(gd["n_burnin"]))
    ch = s.get_chain()
    assert np.array_equal(np.array(ch.vals["accepted"], dtype=np.uint8), gd["accepted"]), "accept decisions differ"
    assert 0.2 < gd["accepted"].mean() < 0.9
    assert rel_err(npy(ch.get_samples()), gd["samples"]) < 1e-9
    assert rel_err(npy(ch.get_target_vals()), gd["target_vals"]) < 1e-9
    assert rel_err(npy(ch.get_grad_vals()), gd["grad_vals"]) < 1e-8
    assert rel_err(npy(s.current["sample"]), gd["final_sample"]) < 1e-9


@pytest.mark.parametrize("arch,data,C,T,step", [("2321", "noisy", 37, 10, 0.6), ("2321", "xor", 20, 25, 1.0),
                                                ("221", "xor", 11, 40, 1.2), ("221", "noisy", 9, 12, 0.5)])
def test_smmala_tape_vs_oracle(arch, data, C, T, step):
    rng = np.random.default_rng(C)
    spec = spec_of(arch)
    P = spec.num_params
    x, y = noisy_xor() if data == "noisy" else data_of(arch, np.float64)
    theta0 = rng.normal(size=(C, P)) * 0.8
    z, u = rng.normal(size=(T, C, P)), rng.uniform(size=(T, C))
    nb = 2
    ref = oracle.smmala_run(spec, x, y, np.zeros(P), np.full(P, S3), theta0, z, u, step, n_burnin=nb)
    m = make_model(arch, "f64", S3)
    ds = XYDataset(torch.from_numpy(x), torch.from_numpy(y))
    s = SMMALA(m, theta0=torch.from_numpy(theta0), dataloader=loader(ds), step=step)
    s.set_noise_tape(torch.from_numpy(z), torch.from_numpy(u))
    s.run(num_epochs=T, num_burnin_epochs=nb)
    got = s.get_chain()
    assert 0.05 < ref["accepted"].mean() < 0.999
    assert np.array_equal(npy(got.accepted_soa), ref["accepted"])
    assert rel_err(npy(got.get_samples().permute(1, 0, 2)), ref["sample"]) < 1e-9
    assert rel_err(npy(got.target_soa), ref["target_val"]) < 1e-9
    assert rel_err(npy(s.current["grad_val"]), ref["final"]["grad_val"]) < 1e-8


def test_smmala_single_chain_api_and_philox():
    arch, P, T, step, seed = "2321", 20, 15, 0.8, 77
    x, y = noisy_xor()
    m = make_model(arch, "f64", S3)
    ds = XYDataset(torch.from_numpy(x), torch.from_numpy(y))
    theta0 = np.random.default_rng(1).normal(size=P) * 0.5
    s = SMMALA(m, theta0=torch.from_numpy(theta0), dataloader=loader(ds), step=step, seed=seed)
    s.run(num_epochs=T, num_burnin_epochs=0)
    ch = s.get_chain()
    zz = np.stack([oracle.chain_normals(seed, [0], t, P) for t in range(T)])
    uu = np.stack([oracle.chain_uniforms(seed, [0], t) for t in range(T)])
    ref = oracle.smmala_run(spec_of(arch), x, y, np.zeros(P), np.full(P, S3), theta0[None], zz, uu, step)
    assert ch.vals["accepted"] == ref["accepted"][:, 0].tolist()
    assert rel_err(npy(ch.get_samples()), ref["sample"][:, 0]) < 1e-8
    assert len(ch) == T


def test_smmala_rejects_multiclass_networks():
    m = make_model("433", "f64", S3)
    ds = dataset("433", "f64")
    with pytest.raises(ValueError, match="binary"):
        s = SMMALA(m, theta0=torch.zeros(27, dtype=torch.float64), dataloader=loader(ds), step=0.1)
        s.run(num_epochs=2, num_burnin_epochs=0)


def test_smmala_config3_size_properties():
    """BASELINE config 3 size: 16,384 chains, N=200.  Stored targets equal fresh evaluations at the stored samples."""
    x, y = noisy_xor()
    m = make_model("2321", "f64", S3)
    ds = XYDataset(torch.from_numpy(x), torch.from_numpy(y))
    C = 16384
    theta0 = torch.randn(C, 20, dtype=torch.float64, generator=torch.Generator().manual_seed(0)) * 0.5
    s = SMMALA(m, theta0=theta0, dataloader=loader(ds), step=0.7, seed=5)
    s.run(num_epochs=4, num_burnin_epochs=0)
    got = s.get_chain()
    lt = m.log_target_batch(got.get_samples()[:, -1].contiguous(), ds.x, ds.y)
    assert torch.allclose(lt, got.target_soa[-1], rtol=1e-11, atol=0)
    acc = got.acceptance().mean().item()
    assert 0.02 < acc < 1.0
