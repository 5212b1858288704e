from .am import AM
from .data_sharded_hmc import DataShardedHMC, shard_rows
from .hmc import HMC
from .mala import MALA
from .metropolis_hastings import MetropolisHastings
from .power_posterior_sampler import PowerPosteriorSampler
from .ram import RAM
from .sampler import Sampler
from .serial_sampler import SerialSampler
from .native import NativeChainSampler

# the reference's accessor base class (eeyore/samplers/single_chain_serial_sampler.py) is part of the native sampler base
SingleChainSerialSampler = NativeChainSampler
from .smmala import SMMALA
