"""The two entry points every sampler of the package answers to (API parity with eeyore/samplers/sampler.py:1-8).

`draw` is one transition on the batch (x, y); `run` is the epoch loop.  The native samplers implement `run` as ONE fused
kernel launch when the data loader holds a single batch (samplers/native.py) and fall back to `draw` per mini-batch otherwise.
"""
import abc


class Sampler(abc.ABC):
    def draw(self, x, y, savestate=False):
        """Advance the chain(s) by one iteration on the data (x, y); append the new state to the chain if `savestate`.
        Samplers that only exist as whole runs (the tempered ensemble) do not offer it."""
        raise NotImplementedError(f"{type(self).__name__} has no per-iteration draw()")

    @abc.abstractmethod
    def run(self, num_epochs, num_burnin_epochs, verbose=False, verbose_step=100):
        """`num_epochs` more passes over the data loader; states from iteration `num_burnin_epochs * num_batches` on are kept."""
