"""Flat-parameter model base; mirror of eeyore/models/model.py:5-55.

The reference keeps parameters inside nn.Linear modules and re-views them into the caller's theta on every
``set_params``.  Here the model owns ONE flat device tensor in the reference's layout (per layer: W row-major
[out, in], then bias) and the native kernels read it directly.
"""
import hashlib

import torch

from .. import _native as nv


class Model:
    def __init__(self, dtype=torch.float64, device=None):
        if dtype not in nv.DTYPE_IDS:
            raise ValueError(f"dtype must be torch.float32 or torch.float64, got {dtype}")
        if device is not None and torch.device(device).type != "cuda":
            raise RuntimeError("eeyore_b200 evaluates on a CUDA device only (no CPU path); pass device='cuda[:i]' "
                               "or leave device=None.  Host tensors are accepted as inputs and copied over.")
        self.dtype = dtype
        self._device_arg = device
        self._theta = None

    @property
    def device(self):
        if self._device_arg is not None:
            return torch.device(self._device_arg)
        return torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)

    def _to_dev(self, t):
        """Host or device tensor -> contiguous device tensor of the model dtype."""
        nv.require_cuda()
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t)
        return t.detach().to(device=self.device, dtype=self.dtype, non_blocking=True).contiguous()

    def num_params(self):
        raise NotImplementedError

    def get_params(self):
        return self._theta

    def set_params(self, theta, grad_val=None):
        """model.py:44-55.  ``grad_val`` is kept for signature compatibility and stored as the cached gradient."""
        theta = self._to_dev(theta).reshape(-1)
        if theta.numel() < self.num_params():
            raise ValueError(f"theta has {theta.numel()} entries, the model has {self.num_params()} parameters")
        # like the reference, extra trailing entries are ignored (tests/test_binary_classif_mlp2321_log_lik.py:19-22)
        self._theta = theta[: self.num_params()]
        self._grad_cache = None if grad_val is None else self._to_dev(grad_val).reshape(-1)

    def summary(self, hashsummary=False):
        print(self)
        print("-" * 80)
        print(f"Number of model parameters: {self.num_params()}")
        print("-" * 80)
        if hashsummary:
            print("Hash Summary:")
            for idx, h in enumerate(self.hashsummary()):
                print(f"{idx}: {h}")

    def hashsummary(self):
        return [hashlib.sha256(p.detach().cpu().numpy().tobytes()).hexdigest() for p in self.parameters()]
