"""HMC for ONE chain whose data set is sharded over the GPUs of a box (BASELINE config 5).

Every rank holds the same chain state (theta, momentum, Philox stream) and its own row shard.  One log-target
evaluation = local fused forward+backward kernel over the shard (eeyore_b200/csrc/datapar.cu) -> all-reduce of
1 + P partial sums in fp64 over NCCL -> prior added once; the leapfrog update and the accept test are then computed
redundantly (and identically) on every rank, so no further exchange is needed.
Mirrors eeyore/samplers/hmc.py:100-170 (leapfrog, hamiltonian, linear-space accept); device code: dp_hmc_* kernels.
"""
import ctypes as C

import torch

from .. import _native as nv
from ..chains import ChainList


def shard_rows(n_rows, world_size, rank, multiple=4):
    """Contiguous row range [lo, hi) of `rank`; boundaries are multiples of `multiple` (16-byte aligned fp32 y)."""
    per = -(-n_rows // world_size)
    per = -(-per // multiple) * multiple
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


class DataShardedHMC:
    def __init__(self, model, theta0, x_shard, y_shard, step=0.1, num_steps=10, group=None, seed=0, chain=None,
                 reduce_fn=None):
        if not model.is_data_parallel():
            raise ValueError("DataShardedHMC serves the data-parallel architecture (MLP 16-64-64-1, float32, binary)")
        nv.require_cuda()
        self.model, self.step, self.num_steps, self.group, self.seed = model, float(step), int(num_steps), group, int(seed)
        self.x = model._to_dev(x_shard)
        self.y = model._to_dev(y_shard).reshape(-1)
        if self.x.data_ptr() % 16 or self.y.data_ptr() % 16:
            raise ValueError("row shards must start at 16-byte aligned addresses")
        self.chain = chain if chain is not None else ChainList(keys=["sample", "target_val", "accepted"])
        self._reduce = reduce_fn or self._all_reduce
        dev, p = self.x.device, model.num_params()
        f32, f64 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.float64, device=dev)
        self._sums = torch.empty(p + 1, **f64)
        self._work = torch.empty(int(nv.lib().eeyore_b200_dp_workspace_bytes()) // 8, **f64)   # per-CTA partial sums
        self._theta_p, self._grad_p, self._mom = (torch.empty(p, **f32) for _ in range(3))
        self._lt_c, self._lt_p, self._kin0, self._kin1 = (torch.empty(1, **f64) for _ in range(4))
        self._acc_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self._iter = 0
        self._tape = None
        self.n_evals = 0
        self._theta_c = torch.empty(p, **f32)
        self._grad_c = torch.empty(p, **f32)
        self.current = {"sample": self._theta_c, "target_val": self._lt_c[0], "grad_val": self._grad_c, "accepted": None}
        self.reset(theta0)

    def reset(self, theta, data=None, reset_counter=True, reset_chain=True):
        """SingleChainSerialSampler.reset (single_chain_serial_sampler.py:33-38): restart the chain at theta (host or
        device tensor) re-using every device buffer; re-evaluates target and gradient there."""
        self._theta_c.copy_(torch.as_tensor(theta).reshape(-1), non_blocking=True)
        if reset_chain:
            self.chain.reset(keys=list(self.chain.vals.keys()))
        if reset_counter:
            self._acc_count.zero_()
        self._evaluate(self._theta_c, self._lt_c, self._grad_c)

    def _all_reduce(self, t):
        if torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size(self.group) > 1:
            torch.distributed.all_reduce(t, group=self.group)

    def _evaluate(self, theta, out_target, out_grad):
        m, lib = self.model, nv.lib()
        loc, scale = m.prior_on_device()
        st = nv.stream_ptr(self.x.device)
        nv.check(lib.eeyore_b200_dp_loglik_grad(nv.ptr(theta), nv.ptr(self.x), nv.ptr(self.y), self.x.shape[0],
                                                nv.ptr(self._sums), nv.ptr(self._work), st))
        self._reduce(self._sums)                 # the one exchange step of the path: 1 + P doubles
        nv.check(lib.eeyore_b200_dp_finish(nv.ptr(self._sums), nv.ptr(theta), nv.ptr(loc), nv.ptr(scale),
                                           0 if m.temperature is None else 1,
                                           0.0 if m.temperature is None else float(m.temperature),
                                           nv.ptr(out_target), nv.ptr(out_grad), st))
        self.n_evals += 1

    def set_noise_tape(self, z, u):
        m = self.model
        self._tape = [m._to_dev(z).reshape(-1, m.num_params()), m._to_dev(u).reshape(-1), 0]

    def draw(self, out_sample=None, out_target=None, out_acc=None):
        """One HMC iteration (hmc.py:126-170), enqueued without any host synchronisation."""
        lib, st = nv.lib(), nv.stream_ptr(self.x.device)
        zt = ut = None
        if self._tape is not None:
            z, u, pos = self._tape
            zt, ut = z[pos], u[pos:pos + 1]
            self._tape[2] = pos + 1
        with torch.cuda.device(self.x.device):
            nv.check(lib.eeyore_b200_dp_hmc_begin(nv.ptr(self._theta_c), nv.ptr(self._grad_c), self.step, self.seed,
                                                  self._iter, nv.ptr(zt), nv.ptr(self._mom), nv.ptr(self._theta_p),
                                                  nv.ptr(self._kin0), st))
            for s in range(self.num_steps):
                self._evaluate(self._theta_p, self._lt_p, self._grad_p)
                nv.check(lib.eeyore_b200_dp_hmc_step(nv.ptr(self._grad_p), self.step, 1 if s == self.num_steps - 1 else 0,
                                                     nv.ptr(self._mom), nv.ptr(self._theta_p), nv.ptr(self._kin1), st))
            nv.check(lib.eeyore_b200_dp_hmc_accept(nv.ptr(self._theta_c), nv.ptr(self._grad_c), nv.ptr(self._lt_c),
                                                   nv.ptr(self._theta_p), nv.ptr(self._grad_p), nv.ptr(self._lt_p),
                                                   nv.ptr(self._kin0), nv.ptr(self._kin1), self.seed, self._iter, nv.ptr(ut),
                                                   nv.ptr(out_sample), nv.ptr(out_target), nv.ptr(out_acc),
                                                   nv.ptr(self._acc_count), st))
        self._iter += 1

    def run(self, num_epochs, num_burnin_epochs, verbose=False, verbose_step=100):
        """sampler.run of the reference (serial_sampler.py:35-52) with one full-data batch per epoch."""
        dev, p = self.x.device, self.model.num_params()
        n_saved = max(0, num_epochs - num_burnin_epochs)
        samples = torch.empty(n_saved, p, dtype=torch.float32, device=dev)
        targets = torch.empty(n_saved, dtype=torch.float64, device=dev)
        accepted = torch.empty(n_saved, dtype=torch.uint8, device=dev)
        for t in range(num_epochs):
            k = t - num_burnin_epochs
            if k >= 0:
                self.draw(samples[k], targets[k:k + 1], accepted[k:k + 1])
            else:
                self.draw()
        if n_saved:
            self.chain.extend_from_device(samples=samples, target_vals=targets.to(torch.float32), accepted=accepted)
            self.current["accepted"] = int(accepted[-1].item())
        self.current["target_val"] = self._lt_c[0]
        return samples, targets, accepted

    def get_chain(self):
        return self.chain

    def acceptance_count(self):
        return int(self._acc_count.item())
