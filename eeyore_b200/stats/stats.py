"""Chain diagnostics on the device; mirror of eeyore/stats/{cov,inse_mc_cov,mc_cov,mc_se,multi_ess,running_mean}.py.

The reference computes the INSE estimator with a python double loop of torch.ger outer products (0.1-0.7 s per
chain); here one kernel launch (eeyore_b200/csrc/stats.cu) handles all chains, one CTA per chain with the chain
staged in shared memory.  `acf` is builder-defined (it lives in the absent `kanga` package).
"""
import torch

from .. import _native as nv


def running_mean(x, dim=-1):
    """Mirror of eeyore/stats/running_mean.py: cumulative mean along `dim`."""
    n = torch.arange(1, x.shape[dim] + 1, dtype=x.dtype, device=x.device)
    shape = [1] * x.dim()
    shape[dim] = -1
    return torch.cumsum(x, dim=dim) / n.view(shape)


def _device_of(x):
    nv.require_cuda()
    return x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())


DEFER_AFTER = 15     # lag pairs a chain may take in the first launch before it is left to the second one


def chain_stats(x, layout="cnp", want=("mean", "cov", "inse", "ess"), max_lag=None, check=True, defer=True):
    """Diagnostics of C chains in one launch.

    x: [C, n, P] (layout 'cnp', the ChainLists.get_samples orientation) or [n, P, C] (layout 'npc', the samplers'
    device storage); float32/float64, host or device (host tensors are copied over).  Returns a dict of device tensors.
    """
    if x.dtype not in nv.DTYPE_IDS:
        raise ValueError("x must be float32 or float64")
    dev = _device_of(x)
    x = x.to(dev)
    if layout == "cnp":
        c, n, p = x.shape
        ss_chain, ss_iter, ss_param = x.stride()
    elif layout == "npc":
        n, p, c = x.shape
        ss_iter, ss_param, ss_chain = x.stride()
    else:
        raise ValueError(layout)
    if n < 2:
        raise RuntimeError("Not enough samples")
    mk = lambda *shape: torch.empty(*shape, dtype=x.dtype, device=dev)
    out = {}
    if "mean" in want:
        out["mean"] = mk(c, p)
    if "cov" in want:
        out["cov"] = mk(c, p, p)
    if "inse" in want:
        out["inse"] = mk(c, p, p)
    if "ess" in want:
        out["ess"] = mk(c)
    need_status = "inse" in want or "ess" in want
    status = torch.zeros(c, dtype=torch.int32, device=dev)
    lags = torch.zeros(c, 2, dtype=torch.int32, device=dev)
    if max_lag is not None:
        out["acf"] = mk(c, max_lag + 1, p)
    def launch(count, index, defer_after):
        nv.check(nv.lib().eeyore_b200_chain_stats(
            nv.DTYPE_IDS[x.dtype], count, n, p, nv.ptr(x), ss_iter, ss_chain, ss_param,
            nv.ptr(out.get("mean")), nv.ptr(out.get("cov")), nv.ptr(out.get("inse")), nv.ptr(out.get("ess")),
            nv.ptr(status), nv.ptr(lags), -1 if max_lag is None else int(max_lag), nv.ptr(out.get("acf")),
            nv.ptr(index), defer_after, nv.stream_ptr(dev)))

    # A chain whose INSE estimate never becomes positive definite walks through n / 2 lag pairs (and ends as 'Not enough
    # samples'), a hundred times the work of an ordinary chain.  With many chains in the batch the first launch leaves such
    # chains undecided (status 3) after DEFER_AFTER lag pairs, and a second launch finishes all of them side by side, one
    # warp each, instead of each one holding up the warp that happened to meet it.
    # defer = "leave": first launch only; the caller collects the status-3 chains (of several batches) and finishes them in one
    # call with defer=False.
    two_tier = bool(defer) and need_status and (c >= 4096 or defer == "leave")
    with torch.cuda.device(dev):
        launch(c, None, DEFER_AFTER if two_tier else -1)
        if two_tier and defer != "leave":
            todo = (status == 3).nonzero().flatten()
            if todo.numel():
                launch(int(todo.numel()), todo, -1)
    out["status"], out["lags"] = status, lags
    if check and need_status and bool((status == 1).any()):
        raise RuntimeError("Not enough samples")      # inse_mc_cov.py:44-45
    return out


def cov(x, rowvar=False):
    """eeyore/stats/cov.py:5-15 for a [n, P] chain (rowvar=False) or [P, n] (rowvar=True)."""
    if x.dim() > 2:
        raise ValueError("x has more than 2 dimensions")
    if x.dim() < 2:
        x = x.view(1, -1)
    if rowvar or x.size(0) == 1:
        x = x.t()
    return chain_stats(x[None], want=("cov",))["cov"][0]


def inse_mc_cov(x, adjust=False):
    """eeyore/stats/inse_mc_cov.py:9-83.  adjust=True relies on torch.symeig, removed from torch >= 2 (broken in the
    reference as well)."""
    if adjust:
        raise NotImplementedError("adjust=True is not available (the reference path calls the removed torch.symeig)")
    return chain_stats(x[None], want=("inse",))["inse"][0]


def mc_cov(x, method="inse", adjust=False, rowvar=False):
    """eeyore/stats/mc_cov.py."""
    if method == "inse":
        return inse_mc_cov(x, adjust=adjust)
    if method == "iid":
        return cov(x, rowvar=rowvar)
    raise ValueError("The method can be inse or iid, {} was given".format(method))


def mc_cov_batch(x, method="inse", adjust=False):
    """[C, n, P] -> [C, P, P]."""
    if adjust:
        raise NotImplementedError("adjust=True is not available")
    key = {"inse": "inse", "iid": "cov"}[method]
    return chain_stats(x, want=(key,))[key]


def mc_se_from_cov(mc_cov_mat):
    """eeyore/stats/mc_se_from_cov.py: sqrt of the diagonal."""
    return mc_cov_mat.diag().sqrt()


def mc_se(x, method="inse", adjust=False, rowvar=False):
    """eeyore/stats/mc_se.py:4-5: mc_se_from_cov(mc_cov(x)) = sqrt(diag(mc_cov)) -- as in the reference there is no
    division by n (ChainList.mc_se and ChainLists.mc_se give the same value whichever way they are called)."""
    return mc_se_from_cov(mc_cov(x, method=method, adjust=adjust, rowvar=rowvar))


def multi_ess(x, mc_cov_mat=None, method="inse", adjust=False):
    """eeyore/stats/multi_ess.py:6-14 for one [n, P] chain; returns a python float like the reference."""
    if mc_cov_mat is not None or method != "inse":
        n, p = x.shape
        lam = torch.det(cov(x)).item()
        sig = torch.det(mc_cov(x, method=method, adjust=adjust) if mc_cov_mat is None else mc_cov_mat.to(x.dtype)).item()
        return n * ((lam / sig) ** (1 / p))
    if adjust:
        raise NotImplementedError("adjust=True is not available")
    return chain_stats(x[None], want=("ess",))["ess"][0].item()


def multi_ess_batch(x, method="inse", adjust=False, check=True):
    """[C, n, P] -> [C] (ChainLists.multi_ess, chain_lists.py:108-117)."""
    if method != "inse" or adjust:
        raise NotImplementedError
    return chain_stats(x, want=("ess",), check=check)["ess"]


def multi_ess_soa(samples_soa, check=True):
    """The samplers' device storage [n, P, C] -> [C], without any transposition."""
    return chain_stats(samples_soa, layout="npc", want=("ess",), check=check)["ess"]


def acf(x, max_lag):
    """[n, P] -> [max_lag+1, P]; rho_k = sum_t (x_t - mean)(x_{t+k} - mean) / sum_t (x_t - mean)^2 (builder-defined)."""
    return chain_stats(x[None], want=(), max_lag=max_lag)["acf"][0]


def acf_soa(samples_soa, max_lag):
    """[n, P, C] -> [C, max_lag+1, P]."""
    return chain_stats(samples_soa, layout="npc", want=(), max_lag=max_lag)["acf"]


def cor_from_cov(x):
    """eeyore/stats/cor_from_cov.py."""
    sd = x.diag().sqrt()
    return x / sd[:, None] / sd[None, :]


def cor(x, rowvar=False):
    """eeyore/stats/cor.py."""
    return cor_from_cov(cov(x, rowvar=rowvar))


def mc_cor(x, method="inse", adjust=False, rowvar=False):
    """eeyore/stats/mc_cor.py."""
    return cor_from_cov(mc_cov(x, method=method, adjust=adjust, rowvar=rowvar))


def _is_pd(m):
    return bool(torch.equal(m, m.t()) and torch.linalg.cholesky_ex(m).info.item() == 0)


def nearest_pd(a):
    """Higham's nearest positive-definite matrix; mirror of eeyore/linalg/nearest_pd.py:9-42 (the reference's fallback
    loop calls the removed torch.eig; torch.linalg.eigvalsh is used here)."""
    b = (a + a.t()) / 2
    _, s, vh = torch.linalg.svd(b)
    h = vh.t() @ torch.diag(s) @ vh
    a3 = (b + h) / 2
    a3 = (a3 + a3.t()) / 2
    spacing = torch.finfo(a.dtype).eps * torch.linalg.norm(a).item()
    eye, k = torch.eye(a.shape[0], dtype=a.dtype, device=a.device), 1
    while not _is_pd(a3) and k < 100:
        mineig = torch.linalg.eigvalsh(a3).min().item()
        a3 = a3 + eye * (-mineig * k ** 2 + spacing)
        k += 1
    return a3


def multi_rhat(x, mc_cov_mat=None, method="inse", adjust=False):
    """Multivariate potential scale reduction factor; mirror of eeyore/stats/multi_rhat.py:10-40 for x [C, n, P].
    The per-chain INSE covariances and chain means come from one device launch (chain_stats); what is left is P x P
    algebra.  Returns (rhat, imag part, W, B, is_w_pd, is_b_pd) like the reference."""
    c, n, p = x.shape
    if mc_cov_mat is None:
        key = {"inse": "inse", "iid": "cov"}[method]
        if adjust:
            raise NotImplementedError("adjust=True is not available")
        out = chain_stats(x, want=("mean", key))
        w, means = out[key].mean(0), out["mean"]
    else:
        w = torch.stack(list(mc_cov_mat)).to(x.device if x.is_cuda else _device_of(x)).mean(0)
        means = chain_stats(x, want=("mean",))["mean"]
    is_w_pd = _is_pd(w)
    if not is_w_pd:
        w = nearest_pd(w)
    b = chain_stats(means[None], want=("cov",))["cov"][0]          # cov(x.mean(1)), multi_rhat.py:28
    is_b_pd = _is_pd(b)
    if not is_b_pd:
        b = nearest_pd(b)
    eig = torch.linalg.eigvals(torch.linalg.inv(w) @ b)
    i = eig.real.argmax().item()
    rhat = ((n - 1) / n) + ((c + 1) / c) * eig.real[i].item()
    return rhat, eig.imag[i].item(), w, b, is_w_pd, is_b_pd
