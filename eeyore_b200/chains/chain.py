"""Storage interface of a Markov chain: what a sampler needs from the object it appends states to.

API parity with eeyore/chains/chain.py:3-13 (same three method names).  The fused samplers of this package hand over whole
blocks of saved states at once (`extend_from_device` on the concrete classes); `update` / `detach_and_update` remain for the
reference's per-iteration protocol (`sampler.draw(..., savestate=True)`).
"""
import abc

import torch


def _detached(value):
    """A private copy of a tensor that shares neither storage nor autograd history with its source; other values pass."""
    return value.detach().clone() if torch.is_tensor(value) else value


class Chain(abc.ABC):
    @abc.abstractmethod
    def reset(self):
        """Forget every stored state."""

    @abc.abstractmethod
    def update(self, state):
        """Append one state: a dict keyed like the chain ('sample', 'target_val', 'grad_val', 'accepted')."""

    def detach_and_update(self, state):
        self.update({key: _detached(value) for key, value in state.items()})
