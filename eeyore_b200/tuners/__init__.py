from .hmcda_tuner import HMCDATuner
from .tuner import Tuner
