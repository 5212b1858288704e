"""CPU tests of the host-side mirror of the reference interface (no GPU, no compute calls)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from eeyore_b200 import _native as nv
from eeyore_b200.chains import ChainFile, ChainList, ChainLists
from eeyore_b200.constants import loss_functions
from eeyore_b200.datasets import DataCounter, XYDataset
from eeyore_b200.models.mlp import MLP, Hyperparameters

ROOT = Path(__file__).resolve().parent.parent


def test_cabi_library_exports_every_declared_symbol():
    """include/eeyore_b200.h is the contract: every function it declares is exported by the built library and bound
    by the ctypes layer (no compute calls here)."""
    header = (ROOT / "include" / "eeyore_b200.h").read_text()
    declared = set(re.findall(r"\b(eeyore_b200_[a-z0-9_]+)\s*\(", header))
    declared -= {"eeyore_b200_run_params"}
    assert len(declared) >= 14
    assert declared == set(nv.SIGNATURES), declared ^ set(nv.SIGNATURES)
    lib = nv.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.eeyore_b200_version()
    assert lib.eeyore_b200_num_saved(1100, 110, 1) == 990
    assert lib.eeyore_b200_num_saved(10, 3, 3) == 3
    assert lib.eeyore_b200_num_saved(5, 5, 1) == 0


def test_run_params_struct_layout_matches_header():
    """Field order of the ctypes mirror follows the C struct."""
    header = (ROOT / "include" / "eeyore_b200.h").read_text()
    body = header[header.index("typedef struct eeyore_b200_run_params {"):header.index("} eeyore_b200_run_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.split("{")[-1].strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.findall(r"\*?\s*([a-z_0-9]+)\s*(?:\[\d+\])?\s*$", part.strip())[0])
    assert names == [f[0] for f in nv.RunParams._fields_]


def test_mlp_create_errors_without_gpu():
    lib = nv.lib()
    h = C.c_void_p()
    dims = (C.c_int * 3)(2, 2, 1); bias = (C.c_int * 2)(1, 1); acts = (C.c_int * 2)(1, 1)
    assert lib.eeyore_b200_mlp_create(2, dims, bias, acts, 0, 1, C.byref(h)) == 0
    assert lib.eeyore_b200_mlp_num_params(h) == 9
    assert lib.eeyore_b200_mlp_is_specialised(h) == 1
    lib.eeyore_b200_mlp_destroy(h)
    # any other small network is served by the runtime-shape kernels
    dims = (C.c_int * 3)(7, 5, 1)
    bias0 = (C.c_int * 2)(0, 1)
    assert lib.eeyore_b200_mlp_create(2, dims, bias0, acts, 0, 1, C.byref(h)) == 0
    assert lib.eeyore_b200_mlp_num_params(h) == 7 * 5 + 6 and lib.eeyore_b200_mlp_is_specialised(h) == 0
    lib.eeyore_b200_mlp_destroy(h)
    # binary loss on a None head is not a probability: unsupported
    acts_bad = (C.c_int * 2)(1, 0)
    rc = lib.eeyore_b200_mlp_create(2, dims, bias, acts_bad, 0, 1, C.byref(h))
    assert rc == nv.EUNSUPPORTED and b"sigmoid output" in lib.eeyore_b200_last_error()
    with pytest.raises(ValueError):
        nv.check(rc)


def test_hyperparameters_validation():
    """mlp.py:15-19: bare ValueError for short dims or mismatched activations."""
    with pytest.raises(ValueError):
        Hyperparameters(dims=[2, 1], bias=[True], activations=[torch.sigmoid])
    with pytest.raises(ValueError):
        Hyperparameters(dims=[2, 2, 1], activations=[torch.sigmoid])
    hp = Hyperparameters()
    assert hp.dims == [1, 2, 1] and hp.bias == [True, True]


@pytest.mark.parametrize("dims,p", [([2, 2, 1], 9), ([2, 3, 2, 1], 20), ([4, 3, 3], 27), ([4, 3, 2, 3], 32),
                                    ([16, 64, 64, 1], 5313)])
def test_mlp_num_params(dims, p):
    nl = len(dims) - 1
    last = torch.sigmoid if dims[-1] == 1 else None
    loss = loss_functions["binary_classification" if dims[-1] == 1 else "multiclass_classification"]
    m = MLP(loss=loss, hparams=Hyperparameters(dims, nl * [True], (nl - 1) * [torch.sigmoid] + [last]))
    assert m.num_params() == p
    assert m.prior.loc.shape == (p,)


def test_mlp_rejects_unknown_loss_and_cpu_device():
    with pytest.raises(ValueError):
        MLP(loss=lambda x, y: 0, hparams=Hyperparameters([2, 2, 1]))
    with pytest.raises(RuntimeError):
        MLP(loss=loss_functions["binary_classification"], hparams=Hyperparameters([2, 2, 1]), device="cpu")


def test_no_cpu_fallback():
    """Without a CUDA device every compute call fails loudly."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = MLP(loss=loss_functions["binary_classification"], hparams=Hyperparameters([2, 2, 1]))
    xor = XYDataset.from_eeyore("xor")
    with pytest.raises(RuntimeError, match="CUDA"):
        m.log_target(torch.zeros(9, dtype=torch.float64), xor.x, xor.y)


def test_product_does_not_import_oracle():
    for f in (ROOT / "eeyore_b200").rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f
    for f in (ROOT / "eeyore_b200" / "csrc").glob("*.cu*"):
        assert "#include \"../../oracle" not in f.read_text() and "oracle.h" not in f.read_text(), f


def test_data_counter():
    c = DataCounter(4, 4)
    assert c.num_batches == 1
    c.set_epoch_info(1100, 110)
    assert (c.num_iters, c.num_burnin_iters) == (1100, 110)
    c = DataCounter(32, 150)
    assert c.num_batches == 5
    c = DataCounter(32, 150, drop_last=True)
    assert c.num_batches == 4
    c.set_epoch_info(3, 1)
    assert (c.num_iters, c.num_burnin_iters) == (12, 4)
    c.increment_idx(); c.increment_idx(3)
    assert c.idx == 4
    c.reset()
    assert c.idx == 0
    c.set_num_epochs(9)
    assert c.num_epochs == 3


def test_xydataset_bundled():
    xor = XYDataset.from_eeyore("xor", dtype=torch.float64)
    assert xor.x.tolist() == [[0, 0], [0, 1], [1, 0], [1, 1]] and xor.y.tolist() == [[0], [1], [1], [0]]
    iris = XYDataset.from_eeyore("iris", yndmin=1, yonehot=True, dtype=torch.float32)
    assert iris.x.shape == (150, 4) and iris.y.shape == (150, 3) and iris.y.sum().item() == 150
    mg = np.load(ROOT / "tests" / "golden" / "model_goldens.npz")
    assert np.array_equal(iris.x.double().numpy().round(6), mg["iris_x"].round(6))
    assert len(xor) == 4 and xor[1][1].item() == 1


def test_chainlist_and_chainfile_roundtrip(tmp_path):
    ch = ChainList(keys=["sample", "target_val", "grad_val", "accepted"])
    g = torch.Generator().manual_seed(0)
    for i in range(7):
        ch.detach_and_update(dict(sample=torch.randn(5, generator=g, dtype=torch.float64),
                                  target_val=torch.randn((), generator=g, dtype=torch.float64),
                                  grad_val=torch.randn(5, generator=g, dtype=torch.float64), accepted=i % 2))
    assert len(ch) == 7 and ch.num_params() == 5
    assert ch.acceptance_rate() == 3 / 7
    assert ch.get_samples().shape == (7, 5) and ch.get_param(2).shape == (7,)
    ch.extend_from_device(samples=torch.zeros(3, 5, dtype=torch.float64), target_vals=torch.zeros(3, dtype=torch.float64),
                          grad_vals=torch.ones(3, 5, dtype=torch.float64), accepted=torch.tensor([1, 1, 0]))
    assert len(ch) == 10 and ch.vals["accepted"][-3:] == [1, 1, 0]
    ch.to_chainfile(path=tmp_path / "run1", mode="w")
    # file format of chain_file.py:28-45
    lines = (tmp_path / "run1" / "sample.csv").read_text().splitlines()
    assert len(lines) == 10 and len(lines[0].split(",")) == 5 and re.fullmatch(r"-?\d\.\d{18}e[+-]\d\d", lines[0].split(",")[0])
    assert (tmp_path / "run1" / "accepted.csv").read_text().splitlines()[:3] == ["0", "1", "0"]
    back = ChainFile(keys=["sample", "target_val", "grad_val", "accepted"], path=tmp_path / "run1").to_chainlist()
    assert torch.equal(back.get_samples(), ch.get_samples())
    assert torch.equal(back.get_target_vals(), ch.get_target_vals())
    assert back.vals["accepted"] == ch.vals["accepted"]
    # append mode, one update per iteration like the reference
    cf = ChainFile(keys=["sample", "accepted"], path=tmp_path / "run2", mode="a")
    cf.update(dict(sample=torch.ones(2, dtype=torch.float64), accepted=1))
    cf.update(dict(sample=torch.zeros(2, dtype=torch.float64), accepted=0))
    assert (tmp_path / "run2" / "accepted.csv").read_text() == "1\n0\n"
    lists = ChainLists.from_chain_list([ch, back], keys=["sample", "accepted"])
    assert lists.num_chains() == 2 and lists.num_samples() == 10 and lists.get_samples().shape == (2, 10, 5)
    assert lists.acceptance() == [ch.acceptance_rate()] * 2
    st = ch.state(-1)
    assert st["accepted"] == 0 and torch.equal(st["sample"], torch.zeros(5, dtype=torch.float64))


def test_model_deepcopy_gets_its_own_native_handle():
    """The reference deep-copies the model per tempered chain (power_posterior_sampler.py:71-83)."""
    import copy
    m = MLP(loss=loss_functions["binary_classification"], hparams=Hyperparameters([2, 2, 1]), dtype=torch.float64)
    h = m.handle()
    c = copy.deepcopy(m)
    c.temperature = 0.25
    assert m.temperature is None and c.num_params() == 9
    assert c.handle().value != h.value


def test_run_params_validation_happens_before_any_launch():
    """Argument errors of the fused runs are reported (EINVAL) before the first CUDA call: Philox counter words are 32 bits,
    so offsets that would alias streams are refused; lanes_per_chain has five legal widths."""
    lib = nv.lib()
    h = C.c_void_p()
    dims = (C.c_int * 4)(2, 3, 2, 1); bias = (C.c_int * 3)(1, 1, 1); acts = (C.c_int * 3)(1, 1, 1)
    assert lib.eeyore_b200_mlp_create(3, dims, bias, acts, 0, 1, C.byref(h)) == 0
    dummy = C.c_void_p(16)            # never dereferenced: validation fails first

    def params(**kw):
        p = nv.RunParams()
        p.n_chains, p.n_iters, p.n_rows, p.thin, p.step, p.num_steps = 8, 4, 4, 1, 0.1, 3
        for k in ("theta", "target", "grad", "x", "y", "prior_loc", "prior_scale"):
            setattr(p, k, dummy.value)
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    for kw, needle in [(dict(chain_offset=1 << 32), b"2^32"), (dict(chain_offset=(1 << 32) - 4), b"2^32"),
                       (dict(iter_offset=(1 << 32) - 2), b"2^32"), (dict(lanes_per_chain=2), b"lanes_per_chain"),
                       (dict(lanes_per_chain=64), b"lanes_per_chain"), (dict(step=0.0), b"step")]:
        p = params(**kw)
        for entry in (lib.eeyore_b200_hmc_run, lib.eeyore_b200_mala_run, lib.eeyore_b200_mh_run):
            assert entry(h, C.byref(p)) == nv.EINVAL, kw
            assert needle in lib.eeyore_b200_last_error(), (kw, lib.eeyore_b200_last_error())
    p = params(iter_offset=(1 << 32) - 2)
    assert lib.eeyore_b200_smmala_run(h, C.byref(p)) == nv.EINVAL
    # tape mode does not touch the Philox counters: large offsets are fine there (n_iters = 0 returns before any launch)
    p = params(iter_offset=(1 << 40), rng_mode=nv.RNG_TAPE, z_tape=dummy.value, u_tape=dummy.value, n_iters=0)
    assert lib.eeyore_b200_hmc_run(h, C.byref(p)) == 0
    lib.eeyore_b200_mlp_destroy(h)


def test_single_chain_serial_sampler_surface():
    """The accessor surface of eeyore/samplers/single_chain_serial_sampler.py lives on the native sampler base."""
    from eeyore_b200 import samplers
    base = samplers.SingleChainSerialSampler
    for name in ("get_model", "get_chain", "get_param", "get_sample", "set_current", "set_all", "reset", "to_chainfile",
                 "run", "benchmark", "draw"):
        assert callable(getattr(base, name)), name
    assert issubclass(samplers.HMC, base) and issubclass(samplers.MALA, base) and issubclass(base, samplers.SerialSampler)
