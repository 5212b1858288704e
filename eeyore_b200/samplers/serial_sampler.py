"""Epoch x batch driver; mirror of eeyore/samplers/serial_sampler.py:8-126.

`run` keeps the reference's loop for mini-batched data loaders (num_batches != 1).  In the full-batch case
(num_batches == 1, every example and BASELINE config) the whole run -- all iterations, burn-in gating, sample
write-out -- is ONE fused kernel launch (`_run_fused`, implemented by the native samplers).
`benchmark` keeps the reference's directory layout but simulates all chains of a round in one batched launch.
"""
from datetime import timedelta
from pathlib import Path
from timeit import default_timer as timer

from .sampler import Sampler


class SerialSampler(Sampler):
    def __init__(self, counter):
        self.counter = counter

    def run(self, num_epochs, num_burnin_epochs, verbose=False, verbose_step=100):
        self.counter.set_epoch_info(num_epochs, num_burnin_epochs)
        if self.counter.num_batches == 1 and hasattr(self, "_run_fused"):
            start = timer()
            # like the reference's loop, every call performs num_epochs * num_batches draws; counter.idx keeps counting
            # across calls, so a second run() continues the chain (burn-in gating by the global index, :46)
            self._run_fused(self.counter.num_iters)
            if verbose:
                print(f"Iterations {self.counter.idx} out of {self.counter.num_iters} (fused launch), "
                      f"duration {timedelta(seconds=timer() - start)}")
            return
        for epoch in range(self.counter.num_epochs):
            for x, y in self.dataloader:
                start = timer()
                self.draw(x, y, savestate=self.counter.idx >= self.counter.num_burnin_iters)
                if verbose and (self.counter.idx + 1) % verbose_step == 0:
                    print(f"Iteration {self.counter.idx + 1} out of {self.counter.num_iters} "
                          f"(in epoch {epoch + 1} out of {self.counter.num_epochs}), "
                          f"duration {timedelta(seconds=timer() - start)}")
                self.counter.increment_idx()

    def benchmark(self, num_chains, num_epochs, num_burnin_epochs, path, init=None, check_conditions=None,
                  verbose=False, verbose_step=100, print_acceptance=False, print_runtime=True):
        """serial_sampler.py:54-126 with all pending chains simulated side by side on the device.
        Writes run%0Nd/{sample,target_val,accepted,...}.csv + runtime.txt, errors/, run_counts.txt."""
        import torch
        path = Path(path)
        done, unmet, errors = 0, 0, 0
        width = len(str(num_chains))
        while done < num_chains:
            todo = num_chains - done
            try:
                if init is None:
                    theta0 = self.get_model().prior.sample((todo,))
                else:
                    theta0 = torch.stack([init[done + i] for i in range(todo)])
                batch = self._spawn(theta0)
                start = timer()
                batch.run(num_epochs=num_epochs, num_burnin_epochs=num_burnin_epochs, verbose=verbose,
                          verbose_step=verbose_step)
                torch.cuda.synchronize()
                runtime = timer() - start
                chains = batch.get_chain()
                for i in range(todo):
                    chain_i = chains.to_chainlist(i) if hasattr(chains, "to_chainlist") else chains
                    if check_conditions is None or check_conditions(chain_i, runtime):
                        run_path = path / ("run" + str(done + 1).zfill(width))
                        run_path.mkdir(parents=True, exist_ok=True)
                        chain_i.to_chainfile(path=run_path, mode="w")
                        (run_path / "runtime.txt").write_text(f"{runtime}\n")
                        done += 1
                        if verbose:
                            msg = "Succeeded"
                            if print_acceptance:
                                msg += f"; acceptance rate = {chain_i.acceptance_rate()}"
                            if print_runtime:
                                msg += f"; runtime = {timedelta(seconds=runtime)}"
                            print(msg + "\n")
                    else:
                        unmet += 1
            except RuntimeError as error:
                err_path = path / ("run" + str(done + 1).zfill(width)) / "errors"
                err_path.mkdir(parents=True, exist_ok=True)
                (err_path / f"error{str(errors + 1).zfill(width)}.txt").write_text(f"{error}\n")
                errors += 1
                if errors > 10 * num_chains:
                    raise
        path.mkdir(parents=True, exist_ok=True)
        (path / "run_counts.txt").write_text(f"{done},succesful\n{unmet},unmet_conditions\n{errors},runtime_errors\n")
