"""Device diagnostics (multi-ESS, INSE Monte Carlo covariance, covariance, ACF); filled in by stats.py."""
from .stats import *  # noqa: F401,F403
