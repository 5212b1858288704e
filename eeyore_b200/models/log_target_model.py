"""Mirror of eeyore/models/log_target_model.py:7-23."""
import torch

from .model import Model


class LogTargetModel(Model):
    def __init__(self, temperature=None, dtype=torch.float64, device=None):
        super().__init__(dtype=dtype, device=device)
        self.temperature = temperature
        self._grad_cache = None
        self._last_call = None

    def log_target(self, theta, x, y):
        raise NotImplementedError

    def grad_log_target(self, log_target_val=None):
        """log_target_model.py:15-18.  The reference differentiates the autograd scalar it is handed; the native
        model has no autograd graph, so this returns the gradient of the LAST log_target call (computed in the same
        fused kernel as the value).  Deviation documented in DESIGN.md."""
        if self._grad_cache is None:
            if self._last_call is None:
                raise RuntimeError("grad_log_target called before log_target")
            theta, x, y = self._last_call
            _, self._grad_cache = self._eval(theta[None], x, y, want_grad=True)
            self._grad_cache = self._grad_cache[0]
        return self._grad_cache

    def upto_grad_log_target(self, theta, x, y):
        """log_target_model.py:20-23: one fused kernel launch returns both."""
        self.set_params(theta)
        xd, yd = self._to_dev(x), self._to_dev(y)
        lt, g = self._eval(self._theta[None], xd, yd, want_grad=True)
        self._last_call = (self._theta, xd, yd)
        self._grad_cache = g[0]
        return lt[0], g[0]
