"""GPU parity: on-device cov / INSE Monte Carlo covariance / multi-ESS / ACF against the reference's goldens
(examples/stats/chain01..04.csv evaluated by the unmodified reference) and against the oracle."""
import numpy as np
import pytest
import torch

import oracle
from eeyore_b200 import stats as st
from eeyore_b200.chains import ChainList, ChainLists, DeviceChains
from gpu_helpers import npy
from helpers import load, rel_err

pytestmark = pytest.mark.gpu


def test_reference_stats_goldens():
    gd = load("stats_goldens")
    x = torch.from_numpy(gd["chains"])                    # [4, 1000, 3]
    for i in range(4):
        assert rel_err(npy(st.cov(x[i])), gd["cov"][i]) < 1e-12
        assert rel_err(npy(st.inse_mc_cov(x[i])), gd["inse"][i]) < 1e-10
        assert abs(st.multi_ess(x[i]) - gd["multi_ess"][i]) < 1e-9 * gd["multi_ess"][i]
        assert rel_err(npy(st.mc_cov(x[i], method="iid")), gd["cov"][i]) < 1e-12
    ess = npy(st.multi_ess_batch(x))
    assert np.allclose(ess, gd["multi_ess"], rtol=1e-9, atol=0)
    assert abs(ess[0] - 564.6937234344964) < 1e-6            # SURVEY.md section 4 known answer
    lists = ChainLists(vals={"sample": [list(x[i].unbind(0)) for i in range(4)]})
    assert np.allclose(lists.multi_ess(), gd["multi_ess"], rtol=1e-9, atol=0)
    se = npy(st.mc_se(x[0]))                                 # eeyore/stats/mc_se.py: sqrt(diag(mc_cov)), no 1/n
    assert np.allclose(se, np.sqrt(np.diag(gd["inse"][0])), rtol=1e-10)
    ch = ChainList(vals={"sample": list(x[0].unbind(0))})    # the same value along every call path
    assert np.allclose(npy(ch.mc_se()), se, rtol=1e-12)
    assert np.allclose(npy(ch.mc_se(mc_cov_mat=ch.mc_cov())), se, rtol=1e-12)


def test_acf_against_third_party_goldens():
    """ACF over the reference's chain fixtures against scipy.signal.correlate / numpy.correlate (acf_goldens.npz)."""
    gd, ga = load("stats_goldens"), load("acf_goldens")
    x, k = torch.from_numpy(gd["chains"]), int(ga["max_lag"])
    for i in range(4):
        assert np.max(np.abs(npy(st.acf(x[i], k)) - ga["acf"][i])) < 1e-12
    soa = x.permute(1, 2, 0).contiguous()                   # [n, P, C]: the layout the samplers save in
    batch = npy(st.acf_soa(soa.cuda(), k))
    assert np.max(np.abs(batch - ga["acf"])) < 1e-12


def test_config1_chain_multi_ess():
    gd = load("mala_xor221_f64")
    ch = ChainList(vals={"sample": list(torch.from_numpy(gd["samples"]).unbind(0))})
    assert abs(ch.multi_ess() - float(gd["multi_ess"])) < 1e-8 * float(gd["multi_ess"])
    assert rel_err(npy(ch.mc_cov()), oracle.inse_mc_cov(gd["samples"])) < 1e-9


def ar1_chains(C, n, P, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    rho = rng.uniform(0.0, 0.95, size=(C, 1, P))
    mix = rng.normal(size=(C, P, P)) * 0.3 + np.eye(P)
    e = rng.normal(size=(C, n, P))
    x = np.empty((C, n, P))
    x[:, 0] = e[:, 0]
    for t in range(1, n):
        x[:, t] = rho[:, 0] * x[:, t - 1] + e[:, t]
    return np.einsum("cnp,cpq->cnq", x, mix).astype(dtype)


@pytest.mark.parametrize("C,n,P", [(61, 400, 3), (40, 500, 9), (33, 600, 20), (20, 301, 27), (9, 1000, 27), (5, 200, 32),
                                   (7, 64, 1)])
def test_batched_chains_vs_oracle(C, n, P):
    x = ar1_chains(C, n, P, seed=P)
    out = st.chain_stats(torch.from_numpy(x), want=("mean", "cov", "inse", "ess"), max_lag=min(25, n - 1))
    soa = torch.from_numpy(x).cuda().permute(1, 2, 0).contiguous()        # [n, P, C]
    out2 = st.chain_stats(soa, layout="npc", want=("ess", "inse"))
    assert torch.equal(out["ess"], out2["ess"]) and torch.equal(out["inse"], out2["inse"])
    lags = npy(out["lags"])
    for c in range(C):
        sig, info = oracle.inse_mc_cov(x[c], return_info=True)
        assert (info["sn"], info["m_last"]) == tuple(lags[c]), (c, info, lags[c])
        assert rel_err(npy(out["inse"][c]), sig) < 1e-9
        assert rel_err(npy(out["cov"][c]), oracle.cov(x[c])) < 1e-11
        assert abs(out["ess"][c].item() - oracle.multi_ess(x[c])) < 1e-8 * oracle.multi_ess(x[c])
        assert np.allclose(npy(out["mean"][c]), x[c].mean(0), rtol=1e-11, atol=1e-13)
        assert np.max(np.abs(npy(out["acf"][c]) - oracle.acf(x[c], min(25, n - 1)))) < 1e-11


def test_fp32_chains():
    x = ar1_chains(12, 300, 5, seed=1, dtype=np.float32)
    ess = npy(st.multi_ess_batch(torch.from_numpy(x)))
    ref = np.array([oracle.multi_ess(x[c].astype(np.float64)) for c in range(12)])
    assert np.allclose(ess, ref, rtol=2e-2)


def test_not_enough_samples_raises_like_the_reference():
    x = torch.tensor([[0.0, 1.0], [1.0, 0.0], [0.5, 0.5]], dtype=torch.float64)
    with pytest.raises(RuntimeError, match="Not enough samples"):
        st.inse_mc_cov(x)
    with pytest.raises(RuntimeError, match="Not enough samples"):
        st.multi_ess(x)
    out = st.chain_stats(x[None], want=("ess",), check=False)
    assert out["status"].item() == 1 and torch.isnan(out["ess"]).all()


def test_device_chains_diagnostics_after_sampling():
    """Sampler output -> on-device multi-ESS / ACF without leaving the GPU (BASELINE config 4's second half)."""
    from eeyore_b200.samplers import HMC
    from gpu_helpers import dataset, loader, make_model
    m = make_model("2321", "f64", 3 ** 0.5)
    ds = dataset("2321", "f64")
    C, n = 96, 1500
    theta0 = torch.randn(C, 20, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    s = HMC(m, theta0=theta0, dataloader=loader(ds), step=0.3, num_steps=10, seed=11)
    s.run(num_epochs=n + 50, num_burnin_epochs=50)
    chains = s.get_chain()
    assert isinstance(chains, DeviceChains) and chains.num_samples() == n
    out = st.chain_stats(chains.samples_soa, layout="npc", want=("ess",), max_lag=20, check=False)
    ess, acf, status = npy(out["ess"]), npy(out["acf"]), npy(out["status"])
    xs = npy(chains.get_samples())
    for c in range(0, C, 5):
        try:
            ref = oracle.multi_ess(xs[c])
            assert status[c] == 0 and abs(ess[c] - ref) < 1e-7 * abs(ref), c
        except RuntimeError:                      # the reference raises 'Not enough samples' for this chain
            assert status[c] == 1 and np.isnan(ess[c]), c
        assert np.max(np.abs(acf[c] - oracle.acf(xs[c], 20))) < 1e-10
    good = status == 0
    assert good.any() and np.all(ess[good] > 1) and np.allclose(acf[:, 0], 1.0)
    if good.all():
        assert np.allclose(npy(chains.multi_ess()), ess)
    else:
        with pytest.raises(RuntimeError, match="Not enough samples"):
            chains.multi_ess()


def test_multi_rhat_reference_golden():
    """multi_rhat (eeyore/stats/multi_rhat.py:10-40) on the reference's four example chains."""
    gd = load("stats_goldens")
    x = torch.from_numpy(gd["chains"])
    rhat, imag, w, b, w_pd, b_pd = st.multi_rhat(x)
    assert abs(rhat - float(gd["multi_rhat"])) < 1e-9 and abs(rhat - 1.0134832973360262) < 1e-9
    assert imag == 0 and w_pd and b_pd
    lists = ChainLists(vals={"sample": [list(x[i].unbind(0)) for i in range(4)], "accepted": [[1] * 1000] * 4})
    assert abs(lists.multi_rhat()[0] - rhat) < 1e-12
    summ = lists.summary(keys=["mean", "mc_se", "acceptance", "multi_ess", "multi_rhat"])
    assert abs(summ["multi_rhat"] - rhat) < 1e-12 and summ["acceptance"] == 1
    assert abs(summ["multi_ess"] - gd["multi_ess"].mean()) < 1e-6
    assert np.allclose(npy(summ["mc_se"]), np.sqrt(np.stack([np.diag(gd["inse"][i]) for i in range(4)])).mean(0), rtol=1e-9)
    c = npy(st.cor(x[0]))
    assert np.allclose(np.diag(c), 1.0) and np.allclose(c, np.corrcoef(gd["chains"][0].T), rtol=1e-10)
    assert np.allclose(np.diag(npy(st.mc_cor(x[0]))), 1.0)
    m = torch.tensor([[1.0, 2.0], [2.0, 1.0]], dtype=torch.float64)
    from eeyore_b200.stats.stats import _is_pd
    assert _is_pd(st.nearest_pd(m))
