"""eeyore_b200 -- B200-native (sm_100a) implementation of the sampler inner loop of papamarkou/eeyore.

Only the hot path is implemented: MLP log_target + gradient, the MH / MALA / HMC / SMMALA
draws, ChainList / ChainFile / ChainLists output, and multi-ESS / ACF diagnostics.  Module and class names mirror
the reference so that `from eeyore_b200.models import mlp` replaces `from eeyore.models import mlp`.
"""
__version__ = "0.1.0"

from . import _native  # noqa: F401
