"""HMC for ONE chain whose data set is sharded over the GPUs of a box (BASELINE config 5).

Every rank holds the same chain state (theta, momentum, Philox stream) and its own row shard.  One log-target
evaluation = two launches: the tcgen05 forward+backward kernel over the shard (csrc/datapar_tc.cu) and the fused "post"
kernel (csrc/datapar.cu: dp_post_kernel) that folds the per-CTA sums, exchanges the 1 + P fp64 sums with the peers by
direct NVLink stores into their CUDA-IPC-mapped inboxes (sequence-numbered flags, totals added in rank order, so
bit-identical on every rank), adds the prior once and applies the leapfrog update.  The accept test is computed
redundantly (and identically) on every rank, so no further exchange is needed.  exchange="nccl" keeps the plain
all-reduce formulation (also used when a custom reduce_fn simulates the shards on one device).
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
                 reduce_fn=None, exchange="auto", trajectory="persistent"):
        if not model.is_data_parallel():
            raise ValueError("DataShardedHMC serves the data-parallel architecture (MLP 16-64-64-1, float32, binary)")
        nv.require_cuda()
        self.model, self.step, self.num_steps, self.group, self.seed = model, float(step), int(num_steps), group, int(seed)
        self.x = model._to_dev(x_shard)
        self.y = model._to_dev(y_shard).reshape(-1)
        if self.x.data_ptr() % 16 or self.y.data_ptr() % 16:
            raise ValueError("row shards must start at 16-byte aligned addresses")
        if trajectory not in ("persistent", "launches"):
            raise ValueError("trajectory must be 'persistent' (the whole run in one cooperative launch) or 'launches'")
        self.trajectory = trajectory
        self.chain = chain if chain is not None else ChainList(keys=["sample", "target_val", "accepted"])
        self._reduce = reduce_fn or self._all_reduce
        dev, p = self.x.device, model.num_params()
        if self.x.shape[0] < 1:
            raise ValueError("every rank needs at least one row of the data set")
        self._setup_exchange(exchange, reduce_fn is not None, dev)
        self._x_absmax = torch.zeros(1, dtype=torch.float32, device=dev)       # max |x| of the shard, once per data set
        with torch.cuda.device(dev):
            nv.check(nv.lib().eeyore_b200_dp_absmax(nv.ptr(self.x), self.x.numel(), nv.ptr(self._x_absmax), nv.stream_ptr(dev)))
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

    # ---- exchange set-up ---------------------------------------------------------------------------------------------------
    def _dist_world(self):
        d = torch.distributed
        if d.is_available() and d.is_initialized():
            return d.get_world_size(self.group), d.get_rank(self.group)
        return 1, 0

    def _setup_exchange(self, exchange, has_reduce_fn, dev):
        """exchange: "auto" | "p2p" | "local" | "nccl".  auto = nccl when a reduce_fn is given, else the fused path (peer
        stores over NVLink for world > 1, purely local for one rank)."""
        lib = nv.lib()
        world, rank = self._dist_world()
        if exchange == "auto":
            exchange = "nccl" if has_reduce_fn else ("p2p" if world > 1 else "local")
        if exchange not in ("p2p", "local", "nccl"):
            raise ValueError("exchange must be 'auto', 'p2p', 'local' or 'nccl'")
        if exchange == "p2p" and world == 1:
            exchange = "local"
        if exchange == "local":
            world, rank = 1, 0          # this rank's rows are the whole data set, whatever process group exists
        self.exchange, self._world, self._rank = exchange, world, rank
        self._status = torch.zeros(1, dtype=torch.int32, device=dev)
        self._n_parts = int(lib.eeyore_b200_dp_num_parts(self.x.shape[0]))
        self._xbase, self._xopened, self._peers = None, [], None
        if exchange == "p2p":
            with torch.cuda.device(dev):
                base, handle = C.c_void_p(), (C.c_ubyte * 64)()
                nv.check(lib.eeyore_b200_dp_exchange_create(C.byref(base), handle))
                handles = [None] * world
                torch.distributed.all_gather_object(handles, bytes(handle), group=self.group)
                peers = (C.c_void_p * 8)()
                for r in range(world):
                    if r == rank:
                        peers[r] = base.value
                    else:
                        ptr = C.c_void_p()
                        nv.check(lib.eeyore_b200_dp_exchange_open((C.c_ubyte * 64).from_buffer_copy(handles[r]), C.byref(ptr)))
                        peers[r] = ptr.value
                        self._xopened.append(ptr.value)
                torch.distributed.barrier(group=self.group)          # every area is mapped before the first store
            self._xbase, self._peers = base.value, peers
            self._scratch_ptr = C.c_void_p(base.value + int(lib.eeyore_b200_dp_exchange_scratch_offset()))
        elif exchange == "local":
            self._scratch = torch.zeros(int(lib.eeyore_b200_dp_scratch_len()), dtype=torch.float64, device=dev)
            self._scratch_ptr = nv.ptr(self._scratch)
        self._grid_ctr = torch.zeros(1, dtype=torch.int64, device=dev)

    def close(self):
        """Unmap the peers' exchange areas and free this rank's (collective: every rank calls it)."""
        if getattr(self, "_xbase", None) is not None:
            lib = nv.lib()
            torch.cuda.synchronize(self.x.device)
            if torch.distributed.is_initialized():
                torch.distributed.barrier(group=self.group)
            for ptr in self._xopened:
                lib.eeyore_b200_dp_exchange_close(C.c_void_p(ptr))
            if torch.distributed.is_initialized():
                torch.distributed.barrier(group=self.group)
            lib.eeyore_b200_dp_exchange_destroy(C.c_void_p(self._xbase))
            self._xbase, self._xopened = None, []

    def check_status(self):
        """Raises if a peer never delivered its sums (the device code traps after a bounded spin)."""
        if int(self._status.item()) != 0:
            raise RuntimeError("data-sharded exchange timed out waiting for a peer rank")

    def _all_reduce(self, t):
        if torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size(self.group) > 1:
            torch.distributed.all_reduce(t, group=self.group)

    def _evaluate(self, theta, out_target, out_grad, step_mode=0):
        """target and gradient at theta; step_mode 1 / 2 also applies the inner / last leapfrog update (fused path)."""
        m, lib = self.model, nv.lib()
        loc, scale = m.prior_on_device()
        st = nv.stream_ptr(self.x.device)
        has_t, temp = (0, 0.0) if m.temperature is None else (1, float(m.temperature))
        if self.exchange != "nccl":
            nv.check(lib.eeyore_b200_dp_loglik_grad_x(nv.ptr(theta), nv.ptr(self.x), nv.ptr(self.y), self.x.shape[0],
                                                      nv.ptr(self._x_absmax), None, nv.ptr(self._work), st))
            nv.check(lib.eeyore_b200_dp_post(nv.ptr(self._work), self._n_parts, self._world, self._rank, self.n_evals + 1,
                                             self._peers, self._scratch_ptr, nv.ptr(theta), nv.ptr(loc), nv.ptr(scale), has_t, temp,
                                             nv.ptr(out_grad), nv.ptr(out_target), step_mode, self.step, nv.ptr(self._mom),
                                             nv.ptr(self._theta_p), nv.ptr(self._kin1), nv.ptr(self._status), st))
            self.n_evals += 1
            return
        nv.check(lib.eeyore_b200_dp_loglik_grad_x(nv.ptr(theta), nv.ptr(self.x), nv.ptr(self.y), self.x.shape[0],
                                                  nv.ptr(self._x_absmax), nv.ptr(self._sums), nv.ptr(self._work), st))
        self._reduce(self._sums)                 # the one exchange step of the path: 1 + P doubles
        nv.check(lib.eeyore_b200_dp_finish(nv.ptr(self._sums), nv.ptr(theta), nv.ptr(loc), nv.ptr(scale), has_t, temp,
                                           nv.ptr(out_target), nv.ptr(out_grad), st))
        self.n_evals += 1
        if step_mode:
            nv.check(lib.eeyore_b200_dp_hmc_step(nv.ptr(self._grad_p), self.step, 1 if step_mode == 2 else 0,
                                                 nv.ptr(self._mom), nv.ptr(self._theta_p), nv.ptr(self._kin1), st))

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
                self._evaluate(self._theta_p, self._lt_p, self._grad_p, 2 if s == self.num_steps - 1 else 1)
            nv.check(lib.eeyore_b200_dp_hmc_accept(nv.ptr(self._theta_c), nv.ptr(self._grad_c), nv.ptr(self._lt_c),
                                                   nv.ptr(self._theta_p), nv.ptr(self._grad_p), nv.ptr(self._lt_p),
                                                   nv.ptr(self._kin0), nv.ptr(self._kin1), self.seed, self._iter, nv.ptr(ut),
                                                   nv.ptr(out_sample), nv.ptr(out_target), nv.ptr(out_acc),
                                                   nv.ptr(self._acc_count), st))
        self._iter += 1

    @property
    def persistent(self):
        """The whole run as one cooperative launch (csrc/datapar_tc.cu: dp_hmc_run_kernel); the NCCL formulation and
        trajectory='launches' keep one launch per evaluation."""
        return self.trajectory == "persistent" and self.exchange != "nccl"

    def _run_persistent(self, n_iters, n_burnin, samples, targets, accepted):
        m, lib = self.model, nv.lib()
        loc, scale = m.prior_on_device()
        has_t, temp = (0, 0.0) if m.temperature is None else (1, float(m.temperature))
        zt = ut = None
        if self._tape is not None:
            z, u, pos = self._tape
            if pos + n_iters > z.shape[0]:
                raise RuntimeError("noise tape exhausted")
            zt, ut = z[pos:pos + n_iters].contiguous(), u[pos:pos + n_iters].contiguous()
            self._tape[2] = pos + n_iters
        n_saved = samples.shape[0]
        with torch.cuda.device(self.x.device):
            nv.check(lib.eeyore_b200_dp_hmc_run(
                nv.ptr(self.x), nv.ptr(self.y), self.x.shape[0], nv.ptr(self._x_absmax), nv.ptr(self._theta_c), nv.ptr(self._grad_c),
                nv.ptr(self._lt_c), nv.ptr(self._theta_p), nv.ptr(self._grad_p), nv.ptr(self._mom), nv.ptr(self._work),
                self._scratch_ptr, nv.ptr(self._grid_ctr), nv.ptr(self._status), nv.ptr(loc), nv.ptr(scale), has_t, temp,
                self.step, self.num_steps, n_iters, n_burnin, self.seed, self._iter, nv.ptr(zt), nv.ptr(ut),
                nv.ptr(samples) if n_saved else None, nv.ptr(targets) if n_saved else None, nv.ptr(accepted) if n_saved else None,
                nv.ptr(self._acc_count), self._world, self._rank, self.n_evals + 1, self._peers, nv.stream_ptr(self.x.device)))
        self._iter += n_iters
        self.n_evals += n_iters * self.num_steps

    def run(self, num_epochs, num_burnin_epochs, verbose=False, verbose_step=100):
        """sampler.run of the reference (serial_sampler.py:35-52) with one full-data batch per epoch."""
        dev, p = self.x.device, self.model.num_params()
        n_saved = max(0, num_epochs - num_burnin_epochs)
        samples = torch.empty(n_saved, p, dtype=torch.float32, device=dev)
        targets = torch.empty(n_saved, dtype=torch.float64, device=dev)
        accepted = torch.empty(n_saved, dtype=torch.uint8, device=dev)
        if self.persistent and num_epochs > 0:
            self._run_persistent(num_epochs, num_burnin_epochs, samples, targets, accepted)
        else:
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

    @property
    def mode(self):
        if self.persistent:
            return "persistent: the whole run (every evaluation, exchange, leapfrog and accept step) in one cooperative launch"
        return "one launch per evaluation + one fused post launch (fold, exchange, prior, leapfrog)"

    def launches_per_iteration(self):
        """Kernel launches of one HMC iteration.  persistent: none of its own (one launch per run()); otherwise begin + accept
        + (evaluation, post) per leapfrog step; the NCCL formulation adds the all-reduce and separate finish / step kernels."""
        if self.persistent:
            return 0
        return 2 + (2 if self.exchange != "nccl" else 4) * self.num_steps

    def acceptance_count(self):
        return int(self._acc_count.item())
