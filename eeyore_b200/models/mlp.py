"""MLP with cross-entropy log-likelihood and Gaussian log-prior; mirror of eeyore/models/mlp.py:9-50.

The arithmetic runs in libeeyore_b200.so (eeyore_b200/csrc/mlp_static.cuh) -- one fused kernel evaluates
log-likelihood, log-prior, log-target and the gradient for C chains.  Batched entry points
(``log_target_batch`` / ``upto_grad_log_target_batch``) take theta of shape [C, P].
"""
import ctypes as C

import torch
from torch.distributions import Normal

from .. import _native as nv
from ..constants import NativeLoss
from .bayesian_model import BayesianModel


class Hyperparameters:
    """mlp.py:9-19."""

    def __init__(self, dims=[1, 2, 1], bias=None, activations=None):
        self.dims = list(dims)
        self.bias = list(bias) if bias is not None else (len(self.dims) - 1) * [True]
        self.activations = list(activations) if activations is not None else (len(self.dims) - 1) * [torch.sigmoid]
        if len(self.dims) < 3:
            raise ValueError
        if len(self.dims) != len(self.activations) + 1:
            raise ValueError


def _act_id(a):
    if a is None:
        return nv.ACT_NONE
    if a is torch.sigmoid or a is torch.nn.functional.sigmoid or isinstance(a, torch.nn.Sigmoid):
        return nv.ACT_SIGMOID
    raise ValueError(f"unsupported activation {a!r}: the native kernels implement torch.sigmoid and None")


class MLP(BayesianModel):
    def __init__(self, loss, temperature=None, prior=None, hparams=None, savefile=None, dtype=torch.float64,
                 device=None):
        super().__init__(loss, temperature=temperature, dtype=dtype, device=device)
        if not isinstance(loss, NativeLoss):
            raise ValueError("loss must be one of eeyore_b200.constants.loss_functions "
                             "(binary_classification / multiclass_classification)")
        self.hp = hparams if hparams is not None else Hyperparameters()
        self._handle = None
        self._num_params = sum((d + (1 if b else 0)) * o
                               for d, o, b in zip(self.hp.dims[:-1], self.hp.dims[1:], self.hp.bias))
        self._prior = None
        self._prior_dev = None
        self.prior = prior or self.default_prior()
        self._theta = None
        if savefile:
            raise NotImplementedError("savefile (nn.Module state dict) is not part of the native hot path")

    # -- architecture ------------------------------------------------------------------------------------------
    def __repr__(self):
        return f"MLP(dims={self.hp.dims}, loss={self.loss.name}, dtype={self.dtype})"

    def num_params(self):
        return self._num_params

    def num_hidden_layers(self):
        return len(self.hp.dims) - 2

    def parameters(self):
        """Views of the flat parameter vector in nn.Module.parameters() order (W_0, b_0, W_1, b_1, ...)."""
        i = 0
        for d, o, b in zip(self.hp.dims[:-1], self.hp.dims[1:], self.hp.bias):
            yield self._theta[i:i + d * o].view(o, d)
            i += d * o
            if b:
                yield self._theta[i:i + o]
                i += o

    def handle(self):
        """Native network descriptor (created on first use; raises ValueError for architectures that are not built)."""
        if self._handle is None:
            L = len(self.hp.dims) - 1
            dims = (C.c_int * (L + 1))(*self.hp.dims)
            bias = (C.c_int * L)(*[1 if b else 0 for b in self.hp.bias])
            acts = (C.c_int * L)(*[_act_id(a) for a in self.hp.activations])
            h = C.c_void_p()
            nv.check(nv.lib().eeyore_b200_mlp_create(L, dims, bias, acts, self.loss.loss_id, nv.DTYPE_IDS[self.dtype],
                                                     C.byref(h)))
            self._handle = h
        return self._handle

    def __deepcopy__(self, memo):
        """copy.deepcopy(model) as the reference's multi-chain samplers do (power_posterior_sampler.py:71-83): the copy gets
        its own native handle (created lazily) instead of sharing or pickling the ctypes pointer."""
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_handle":
                new.__dict__[k] = None
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def __del__(self):
        try:
            if self._handle is not None:
                nv.lib().eeyore_b200_mlp_destroy(self._handle)
        except Exception:
            pass

    # -- prior -------------------------------------------------------------------------------------------------
    def default_prior(self):
        """mlp.py:31-35: N(0, I).  Built on the host so that the model can be constructed without a GPU."""
        return Normal(torch.zeros(self.num_params(), dtype=self.dtype), torch.ones(self.num_params(), dtype=self.dtype))

    @property
    def prior(self):
        return self._prior

    @prior.setter
    def prior(self, dist):
        if not isinstance(dist, Normal):
            raise ValueError("the native log-prior implements torch.distributions.Normal (vector loc / scale)")
        self._prior = dist
        self._prior_dev = None

    def prior_on_device(self):
        if self._prior_dev is None:
            p = self.num_params()
            loc = torch.as_tensor(self._prior.loc).expand(p)
            scale = torch.as_tensor(self._prior.scale).expand(p)
            self._prior_dev = (self._to_dev(loc), self._to_dev(scale))
        return self._prior_dev

    # -- evaluation --------------------------------------------------------------------------------------------
    def _check_data(self, x, y):
        if x.dim() != 2 or x.shape[1] != self.hp.dims[0]:
            raise ValueError(f"x must be [N, {self.hp.dims[0]}], got {tuple(x.shape)}")
        k = self.hp.dims[-1]
        n = x.shape[0]
        if y.numel() != n * k:
            raise ValueError(f"y must hold {n}x{k} entries, got {tuple(y.shape)}")

    def is_data_parallel(self):
        """True for the wide architecture served by the data-parallel kernel (eeyore_b200/csrc/datapar.cu): one
        parameter vector, rows spread over the SMs (and, with DataShardedHMC, over the GPUs)."""
        return (list(self.hp.dims) == [16, 64, 64, 1] and self.dtype == torch.float32
                and self.loss.loss_id == nv.LOSS_BINARY and all(self.hp.bias))

    def _eval_data_parallel(self, theta, x, y, group=None):
        """theta [C,P] -> (lt [C] float64->dtype, grad [C,P]); rows of (x, y) are this rank's shard."""
        import ctypes as C
        dev = theta.device
        c, p = theta.shape
        loc, scale = self.prior_on_device()
        sums = torch.empty(p + 1, dtype=torch.float64, device=dev)
        lt64 = torch.empty(1, dtype=torch.float64, device=dev)
        lt = torch.empty(c, dtype=self.dtype, device=dev)
        g = torch.empty(c, p, dtype=self.dtype, device=dev)
        yv = y.reshape(-1)
        with torch.cuda.device(dev):
            st = nv.stream_ptr(dev)
            for i in range(c):
                nv.check(nv.lib().eeyore_b200_dp_loglik_grad(nv.ptr(theta[i]), nv.ptr(x), nv.ptr(yv), x.shape[0],
                                                             nv.ptr(sums), None, st))
                if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                         and getattr(self, "data_sharded", False)):
                    torch.distributed.all_reduce(sums, group=group)
                nv.check(nv.lib().eeyore_b200_dp_finish(nv.ptr(sums), nv.ptr(theta[i]), nv.ptr(loc), nv.ptr(scale),
                                                        0 if self.temperature is None else 1,
                                                        0.0 if self.temperature is None else float(self.temperature),
                                                        nv.ptr(lt64), nv.ptr(g[i]), st))
                lt[i] = lt64[0]
        return lt, g

    def _eval(self, theta, x, y, want_grad=True, parts=False, lanes=0):
        """theta [C,P], x [N,d0], y: device tensors of the model dtype.  Returns (lt [C], grad [C,P] | None[, ll, lp])."""
        nv.require_cuda()
        self._check_data(x, y)
        if self.is_data_parallel() and not parts:
            lt, g = self._eval_data_parallel(theta.contiguous(), x, y)
            return lt, (g if want_grad else None)
        c = theta.shape[0]
        loc, scale = self.prior_on_device()
        dev = theta.device
        lt = torch.empty(c, dtype=self.dtype, device=dev)
        g = torch.empty(c, self.num_params(), dtype=self.dtype, device=dev) if want_grad else None
        ll = torch.empty(c, dtype=self.dtype, device=dev) if parts else None
        lp = torch.empty(c, dtype=self.dtype, device=dev) if parts else None
        with torch.cuda.device(dev):
            nv.check(nv.lib().eeyore_b200_log_target_grad(
                self.handle(), c, nv.ptr(theta), nv.ptr(x), nv.ptr(y), x.shape[0], nv.ptr(loc), nv.ptr(scale),
                0 if self.temperature is None else 1, 0.0 if self.temperature is None else float(self.temperature),
                nv.ptr(lt), nv.ptr(g), nv.ptr(ll), nv.ptr(lp), lanes, nv.stream_ptr(dev)))
        return (lt, g, ll, lp) if parts else (lt, g)

    def forward(self, x):
        """mlp.py:45-50 at the current parameters: [N, d_L] probabilities (binary) or logits (multiclass)."""
        return self.forward_batch(self._theta[None], x)[0]

    __call__ = forward

    def forward_batch(self, theta, x):
        nv.require_cuda()
        theta, x = self._to_dev(theta), self._to_dev(x)
        out = torch.empty(theta.shape[0], x.shape[0], self.hp.dims[-1], dtype=self.dtype, device=theta.device)
        with torch.cuda.device(theta.device):
            nv.check(nv.lib().eeyore_b200_forward(self.handle(), theta.shape[0], nv.ptr(theta), nv.ptr(x), x.shape[0],
                                                  nv.ptr(out), nv.stream_ptr(theta.device)))
        return out

    def log_target_batch(self, theta, x, y, lanes=0):
        """log_target of C chains: theta [C,P] -> [C]."""
        lt, _ = self._eval(self._to_dev(theta), self._to_dev(x), self._to_dev(y), want_grad=False, lanes=lanes)
        return lt

    def upto_grad_log_target_batch(self, theta, x, y, lanes=0):
        """(log_target [C], gradient [C,P]) of C chains in one launch."""
        return self._eval(self._to_dev(theta), self._to_dev(x), self._to_dev(y), want_grad=True, lanes=lanes)
