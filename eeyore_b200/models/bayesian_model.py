"""Mirror of eeyore/models/bayesian_model.py:6-67 (log_lik / log_prior / log_target).

predictive_posterior (MCIntegrator) is a SURVEY.md section 8(f) "next" row and is not implemented.
"""
import torch

from .log_target_model import LogTargetModel


class BayesianModel(LogTargetModel):
    def __init__(self, loss, temperature=None, dtype=torch.float64, device=None):
        super().__init__(temperature=temperature, dtype=dtype, device=device)
        self.loss = loss

    def default_prior(self):
        raise NotImplementedError

    def summary(self, hashsummary=False):
        super().summary(hashsummary=False)
        print(f"Prior: {self.prior}")
        print("-" * 80)
        if hashsummary:
            print("Hash Summary:")
            for idx, h in enumerate(self.hashsummary()):
                print(f"{idx}: {h}")

    def log_lik(self, x, y):
        """bayesian_model.py:30-35, at the current parameters."""
        out = self._eval(self._theta[None], self._to_dev(x), self._to_dev(y), want_grad=False, parts=True)
        return out[2][0]

    def set_params_and_log_lik(self, theta, x, y):
        self.set_params(theta)
        return self.log_lik(x, y)

    def set_params_and_lik(self, theta, x, y):
        return torch.exp(self.set_params_and_log_lik(theta, x, y))

    def log_prior(self):
        """bayesian_model.py:46-50, at the current parameters."""
        p = self.num_params()
        # the prior term does not depend on the data; evaluate it with a one-row dummy data set
        x = torch.zeros(1, self.hp.dims[0], dtype=self.dtype, device=self.device)
        y = torch.zeros(1, self.hp.dims[-1], dtype=self.dtype, device=self.device)
        if self.hp.dims[-1] > 1:
            y[0, 0] = 1
        out = self._eval(self._theta[None], x, y, want_grad=False, parts=True)
        return out[3][0]

    def log_target(self, theta, x, y):
        """bayesian_model.py:52-56."""
        self.set_params(theta)
        xd, yd = self._to_dev(x), self._to_dev(y)
        lt, _ = self._eval(self._theta[None], xd, yd, want_grad=False)
        self._last_call = (self._theta, xd, yd)
        self._grad_cache = None
        return lt[0]
