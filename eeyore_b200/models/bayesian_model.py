"""Mirror of eeyore/models/bayesian_model.py:6-67 (log_lik / log_prior / log_target).

predictive_posterior / predictive_posterior_from_dataset evaluate all posterior samples
in one batched launch instead of the reference's python loop over samples (eeyore/integrators/mcintegrator.py:16-63).
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

    # -- posterior predictive (bayesian_model.py:58-67 via integrators/mcintegrator.py:16-63) -------------------------
    def _stack_samples(self, theta):
        if isinstance(theta, torch.Tensor):
            return self._to_dev(theta).reshape(-1, self.num_params())
        return self._to_dev(torch.stack([torch.as_tensor(t) for t in theta]))

    def predictive_posterior(self, theta, x, y):
        """Monte Carlo average over the posterior samples `theta` of the likelihood of (x, y); NaN integrands are dropped
        (mcintegrator.py:24-25).  Returns (integral, number of dropped samples) like MCIntegrator.integrate."""
        th = self._stack_samples(theta)
        ll = self._eval(th, self._to_dev(x), self._to_dev(y), want_grad=False, parts=True)[2]
        lik = torch.exp(ll)
        keep = ~torch.isnan(lik)
        return lik[keep].mean(), int((~keep).sum().item())

    def predictive_posterior_from_dataset(self, theta, dataset, num_points, shuffle=True, verbose=False, verbose_step=1):
        """Per data point posterior-predictive likelihood of its own label, for `num_points` points of `dataset`
        (mcintegrator.py:27-63).  One batched forward pass of all samples over the selected points.
        Returns (integrals [num_points], indices [num_points], numbers of dropped samples [num_points])."""
        th = self._stack_samples(theta)
        n = len(dataset)
        order = torch.randperm(n) if shuffle else torch.arange(n)
        idx = torch.cat([order] * (-(-num_points // n)))[:num_points]
        x, y = self._to_dev(dataset.x[idx]), self._to_dev(dataset.y[idx])
        out = self.forward_batch(th, x)                                   # [S, num_points, d_L]
        if out.shape[2] == 1:                                             # exp(-BCE) of one point, stats/loss.py:2
            p, yy = out[:, :, 0], y.reshape(1, -1)
            lik = torch.exp(torch.log(p) * yy + torch.log(1 - p) * (1 - yy))
        else:                                                             # exp(-CrossEntropy) of one point, constants.py:17
            cls = torch.argmax(y, 1)
            lik = torch.softmax(out, dim=2)[:, torch.arange(len(cls)), cls]
        keep = ~torch.isnan(lik)
        integrals = torch.where(keep, lik, torch.zeros_like(lik)).sum(0) / keep.sum(0)
        return integrals, idx.to(integrals.device), (~keep).sum(0)
