"""Shared machinery of the device samplers: state buffers, RNG bookkeeping and the fused-launch call.

One `eeyore_b200_*_run` call advances C chains by n iterations and writes the saved states straight into
structure-of-arrays device buffers; `current`, `chain` and `counter` are then updated so that the objects look
exactly as after the reference's python loop (eeyore/samplers/serial_sampler.py:35-52).
"""
import ctypes as C
from pathlib import Path

import torch

from .. import _native as nv
from ..chains import ChainFile, ChainList, DeviceChains
from ..datasets import DataCounter
from .serial_sampler import SerialSampler


class NativeChainSampler(SerialSampler):
    """One sampler object = the state of C >= 1 chains on the device.  It also carries the accessor surface the reference
    keeps in SingleChainSerialSampler (eeyore/samplers/single_chain_serial_sampler.py:5-41: get_model / get_chain /
    get_param / get_sample / set_all / reset / to_chainfile); `eeyore_b200.samplers.SingleChainSerialSampler` is this
    class."""
    _entry = None          # name of the C entry point
    _uses_grad = True

    def _init_native(self, model, theta0, dataloader, data0, counter, chain, seed, lanes_per_chain, thin):
        SerialSampler.__init__(self, counter or DataCounter.from_dataloader(dataloader))
        self.model = model
        self.dataloader = dataloader
        self.thin = thin
        self.lanes_per_chain = lanes_per_chain
        self._user_chain = chain
        self.chain = chain if chain is not None else ChainList(keys=self._default_chain_keys())
        self.seed = int(seed) if seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
        self._iter_offset = 0
        self._tape = None
        self._data_dev = None
        self._device_blocks = []
        # Storage order of the saved states of a batched run: "npc" = [n_saved, P, C] (chain-minor: the sampler's stores are
        # coalesced) or "cnp" = [C, n_saved, P] (every chain contiguous: what the diagnostics kernel streams with bulk
        # copies).  Either way DeviceChains sees the [n, P, C] indexing (a strided view for "cnp").
        self.sample_layout = "npc"
        # Where a batched run saves its states.  False: device buffers (DeviceChains, on-device diagnostics).  True: pinned
        # host buffers -- with unified addressing cudaHostAlloc'd memory is directly addressable by the device, so the kernel's
        # coalesced stores of the saved states ARE the device->host transfer (posted PCIe writes under the computation; no
        # staging copy, no device memory for the samples).  The buffers are kept and re-used by later runs of the same
        # shape; their contents are valid once the stream has been synchronised.
        self.host_output = False
        self._host_blocks = {}
        # With host_output, HMC / MALA / MH on a compiled network specialisation also leave the FINAL state of every run in
        # pinned host memory (`host_current`: 'sample' [C, P] as a view of the chain-minor buffer the kernel stores into,
        # 'target_val' [C], 'accept_count' [C]); None for the other samplers, whose final state is read from `current`.
        self.host_current = None
        self.num_chains = 1
        self.current = {key: None for key in self.keys}
        if theta0 is not None:
            self.set_current(theta0, data=data0)   # set_current makes its own device copy

    def _default_chain_keys(self):
        return ["sample", "target_val", "accepted"]

    # -- accessors (single_chain_serial_sampler.py:10-20,40-41) ---------------------------------------------------
    def get_model(self):
        return self.model

    def get_param(self, idx):
        return self.get_chain().get_param(idx)

    def get_sample(self, idx):
        return self.get_chain().get_sample(idx)

    def set_all(self, theta, data=None):
        return self.set_current(theta, data=data)

    def to_chainfile(self, path=Path.cwd(), mode="a"):
        chain = self.get_chain()
        if isinstance(chain, DeviceChains):
            chain.to_chainfiles(path, mode=mode)       # run%0Nd/<key>.csv, one directory per chain
        else:
            chain.to_chainfile(path=path, mode=mode)

    # -- state ---------------------------------------------------------------------------------------------------
    def _reset_chain(self):
        keys = list(self.chain.vals.keys())
        if isinstance(self.chain, ChainFile):
            self.chain.close()                         # reset() re-opens the files in the chain's own mode
        self.chain.reset(keys=keys)
        self._device_blocks = []

    def _stage(self, x, y):
        m = self.model
        return m._to_dev(x), m._to_dev(y)

    def set_current(self, theta, data=None):
        """<Sampler>.set_current of the reference (e.g. mala.py:26-29): evaluate target (and gradient) at theta.
        theta may be [P] (one chain, reference behaviour) or [C, P] (C independent chains)."""
        m = self.model
        pn = m.num_params()
        batched = theta.dim() == 2
        c = theta.numel() // pn
        # Same number of chains and same data as before (the repeated reset(theta) of a benchmark / restart loop): the
        # device buffers are kept; theta lands in a [C, P] staging buffer (one contiguous, possibly host->device, copy)
        # and the evaluation writes target / gradient in place.
        reuse = (data is None and self._data_dev is not None and getattr(self, "_theta_soa", None) is not None
                 and self.num_chains == c and self._batched == batched)
        if reuse:
            x = y = None
            xd, yd = self._data_dev
            self._stage_cp.copy_(theta.reshape(c, pn), non_blocking=True)
            th = self._stage_cp
            self._acc_count.zero_()
        else:
            x, y = data or next(iter(self.dataloader))
            th = m._to_dev(theta).reshape(c, pn).clone()
            self._batched, self.num_chains = batched, c
            xd, yd = self._stage(x, y)
            self._data_dev = (xd, yd)
            self._stage_cp = th
            self._acc_count = torch.zeros(c, dtype=torch.int32, device=th.device)
        lt, g = m._eval(th, xd, yd, want_grad=self._uses_grad)
        self._set_state(th, lt, g, reuse)
        self._last_accepted = None
        self._on_new_state()
        self._publish_current()
        return x, y

    def _on_new_state(self):
        """Hook: per-chain auxiliary state (tuner, adaptation) is invalid after set_current."""

    def _set_state(self, th, lt, g, reuse=False):
        """Device state in the chain-minor layout ([P, C] storage; `_theta` / `_grad` are its [C, P] views) so that the
        sampler kernels read and write it coalesced."""
        if reuse:
            self._theta_soa.copy_(th.t())
            if g is not None:
                self._grad_soa.copy_(g.t())
            self._lt.copy_(lt)
            return
        self._theta_soa = th.t().contiguous()
        self._grad_soa = g.t().contiguous() if g is not None else None
        self._theta = self._theta_soa.t()
        self._grad = self._grad_soa.t() if g is not None else None
        self._lt = lt

    def _publish_current(self):
        sq = (lambda t: t) if self._batched else (lambda t: t[0])
        self.current["sample"] = sq(self._theta)
        self.current["target_val"] = sq(self._lt)
        if self._uses_grad and "grad_val" in self.current:
            self.current["grad_val"] = sq(self._grad)
        if self._last_accepted is not None:
            self.current["accepted"] = self._last_accepted if self._batched else int(self._last_accepted[0].item())
        m = self.model
        m._theta = self._theta[0].contiguous()  # the reference leaves the model parameters at the current state

    def reset(self, theta, data=None, reset_counter=True, reset_chain=True):
        """single_chain_serial_sampler.py:33-38.  theta may live on the host (pinned memory makes the copy asynchronous)."""
        if reset_counter:
            self.counter.reset()
        if reset_chain:
            self._reset_chain()
        self.set_all(theta.detach(), data=data)

    # -- noise ---------------------------------------------------------------------------------------------------
    def set_noise_tape(self, z, u):
        """Parity mode: feed the reference's proposal noise.  z [T, P] or [T, C, P] standard normals, u [T] or [T, C]
        uniforms, consumed in order by subsequent draw()/run() calls."""
        m = self.model
        z, u = m._to_dev(z), m._to_dev(u)
        self._tape = [z.reshape(z.shape[0], -1, m.num_params()), u.reshape(u.shape[0], -1), 0]

    # -- one fused launch ----------------------------------------------------------------------------------------
    def _fill_params(self, p):
        """Sampler-specific fields of eeyore_b200_run_params."""
        raise NotImplementedError

    def _launch(self, n_iters, n_burnin, xd, yd, want=("sample", "target_val", "accepted")):
        nv.require_cuda()
        m = self.model
        c, pn = self.num_chains, m.num_params()
        dev = self._theta.device
        thin = max(1, int(self.thin))
        n_saved = int(nv.lib().eeyore_b200_num_saved(n_iters, n_burnin, thin))
        out = {}
        chain_major = self.sample_layout == "cnp"
        host = bool(self.host_output) and self._batched

        def alloc(key, shape, dtype):
            if not host:
                return torch.empty(*shape, dtype=dtype, device=dev)
            buf = self._host_blocks.get(key)
            if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
                buf = self._host_blocks[key] = torch.empty(*shape, dtype=dtype).pin_memory()
            return buf

        new_block = (lambda k: alloc(k, (c, n_saved, pn), m.dtype).permute(1, 2, 0)) if chain_major else \
            (lambda k: alloc(k, (n_saved, pn, c), m.dtype))
        if n_saved > 0:
            if "sample" in want:
                out["sample"] = new_block("sample")
            if "grad_val" in want and self._uses_grad:
                out["grad_val"] = new_block("grad_val")
            if "target_val" in want:
                out["target_val"] = alloc("target_val", (n_saved, c), m.dtype)
            out["accepted"] = alloc("accepted", (n_saved, c), torch.uint8)
        loc, scale = m.prior_on_device()
        p = nv.RunParams()
        p.n_chains, p.n_iters, p.n_burnin, p.thin = c, n_iters, n_burnin, thin
        p.has_temperature = 0 if m.temperature is None else 1
        p.temperature = 0.0 if m.temperature is None else float(m.temperature)
        p.seed, p.iter_offset, p.chain_offset = self.seed, self._iter_offset, getattr(self, "chain_offset", 0)
        keep = []
        if self._tape is not None:
            z, u, pos = self._tape
            if pos + n_iters > z.shape[0]:
                raise RuntimeError("noise tape exhausted")
            zz, uu = z[pos:pos + n_iters].contiguous(), u[pos:pos + n_iters].contiguous()
            keep += [zz, uu]
            p.rng_mode, p.z_tape, p.u_tape = nv.RNG_TAPE, zz.data_ptr(), uu.data_ptr()
            self._tape[2] = pos + n_iters
        else:
            p.rng_mode = nv.RNG_PHILOX
        p.x, p.y, p.n_rows = xd.data_ptr(), yd.data_ptr(), xd.shape[0]
        m._check_data(xd, yd)
        p.prior_loc, p.prior_scale = loc.data_ptr(), scale.data_ptr()
        p.theta, p.target = self._theta_soa.data_ptr(), self._lt.data_ptr()
        p.grad = self._grad_soa.data_ptr() if self._uses_grad else None
        p.st_chain, p.st_param = 1, c
        if "sample" in out or "grad_val" in out:
            p.ss_iter, p.ss_param, p.ss_chain = (out.get("sample", out.get("grad_val"))).stride()
        if "sample" in out:
            p.out_samples = out["sample"].data_ptr()
        if "grad_val" in out:
            p.out_grad = out["grad_val"].data_ptr()
        if "target_val" in out:
            p.out_target = out["target_val"].data_ptr()
        if "accepted" in out:
            p.out_accepted = out["accepted"].data_ptr()
        p.accept_count = self._acc_count.data_ptr()
        if host and self._entry in ("eeyore_b200_hmc_run", "eeyore_b200_mala_run", "eeyore_b200_mh_run") and \
                nv.lib().eeyore_b200_mlp_is_specialised(m.handle()) == 1:
            fin = alloc("final_theta", (pn, c), m.dtype)             # chain-minor: a warp's 32 chains store 256 contiguous bytes
            self.host_current = {"sample": fin.t(), "target_val": alloc("final_target", (c,), m.dtype),
                                 "accept_count": alloc("final_accept_count", (c,), torch.int32)}
            p.final_theta, p.fs_chain, p.fs_param = fin.data_ptr(), 1, c
            p.final_target = self.host_current["target_val"].data_ptr()
            p.final_accept_count = self.host_current["accept_count"].data_ptr()
        else:
            self.host_current = None
        p.lanes_per_chain = int(self.lanes_per_chain or 0)
        p.stream = torch.cuda.current_stream(dev).cuda_stream
        self._fill_params(p)
        with torch.cuda.device(dev):
            nv.check(getattr(nv.lib(), self._entry)(m.handle(), C.byref(p)))
        self._iter_offset += n_iters
        return out

    def _wanted_keys(self):
        if isinstance(self.chain, (ChainList, ChainFile)):
            return tuple(self.chain.vals.keys())
        return ("sample", "target_val", "accepted")

    def _store(self, out):
        if not out:
            return
        if self._batched:
            self._device_blocks.append(out)
        else:   # ChainList keeps the block in memory, ChainFile appends it to its <key>.csv files
            self.chain.extend_from_device(
                samples=out["sample"][:, :, 0] if "sample" in out else None,
                target_vals=out["target_val"][:, 0] if "target_val" in out else None,
                grad_vals=out["grad_val"][:, :, 0] if "grad_val" in out else None,
                accepted=out["accepted"][:, 0])

    def get_chain(self):
        """ChainList for one chain (reference behaviour); DeviceChains when the sampler runs C chains."""
        if not getattr(self, "_batched", False):
            return self.chain
        if not self._device_blocks:
            raise RuntimeError("no saved states yet")
        blocks = self._device_blocks
        cat = lambda k: (None if k not in blocks[0] else blocks[0][k] if len(blocks) == 1 else torch.cat([b[k] for b in blocks]))
        return DeviceChains(cat("sample"), cat("target_val"), cat("grad_val"), cat("accepted"),
                            accept_count=self._acc_count, n_iters=self._iter_offset)

    def _run_fused(self, n_iters):
        """Full-batch run: the n_iters iterations of one run() call in one launch (serial_sampler.py:35-52).  As in the
        reference the counter is not rewound by run(): a second call continues the chain, and only iterations whose global
        index is below num_burnin_iters are dropped."""
        if n_iters <= 0:
            return
        xd, yd = self._data_dev
        n_burnin = max(0, min(n_iters, (self.counter.num_burnin_iters or 0) - self.counter.idx))
        out = self._launch(n_iters, n_burnin, xd, yd, want=self._wanted_keys())
        self._store(out)
        if "accepted" in out:   # host_output: a view of the pinned buffer (valid once the stream has been synchronised)
            self._last_accepted = out["accepted"][-1] if out["accepted"].device.type == "cpu" else out["accepted"][-1].to(torch.int64)
        self.counter.increment_idx(n_iters)
        self._publish_current()

    def draw(self, x, y, savestate=False):
        """One iteration on the batch (x, y) -- the reference's per-iteration entry point."""
        xd, yd = self._stage(x, y)
        if self.counter.num_batches != 1:  # e.g. mala.py:49-51: re-evaluate the current state on this mini-batch
            lt, g = self.model._eval(self._theta.contiguous(), xd, yd, want_grad=self._uses_grad)
            self._set_state(self._theta.contiguous(), lt, g)
        out = self._launch(1, 0, xd, yd, want=self._wanted_keys())
        self._last_accepted = out["accepted"][-1].to(torch.int64)
        if savestate:
            self._store(out)
        self._publish_current()

    def acceptance_counts(self):
        """Accepted proposals per chain over every iteration run so far (burn-in included)."""
        return self._acc_count

    def _spawn(self, theta0):
        raise NotImplementedError
