"""CPU oracle for the eeyore sampler inner loop -- TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the reference's algorithm for the hot
path (MLP log_target + gradient, MH / MALA / HMC / SMMALA draws, INSE
multi-ESS, ACF).  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``eeyore_b200/`` imports it and
the product path has no CPU fallback.

Parity status
-------------
* mlp / mh / mala / hmc / am / ram / power posterior / cov / inse_mc_cov / multi_ess: **pinned** -- checked
  against golden vectors generated from the unmodified reference
  (``/root/reference``) by ``oracle/make_golden.py`` and committed under
  ``tests/golden/`` (see ``tests/test_oracle_golden.py``).
* smmala: the reference snapshot has no SMMALA sampler; pinned by tests/golden/smmala_*.npz, runs assembled from the
  reference's own pieces only (make_golden.py:smmala_goldens) -- identical accept decisions, states < 1e-14
* acf: **parity unpinned** -- the reference snapshot does not contain it
  (SURVEY.md section 0); the restatement follows SURVEY.md A.7 / A.10 and is
  builder-defined.

Every function cites the reference file:line it follows (paths relative to
``/root/reference``).
"""

from .mlp import MLPSpec, log_lik, log_prior, log_target, log_target_grad, forward  # noqa: F401
from .samplers import mh_run, mala_run, hmc_run, smmala_run, fisher_metric, DATuner, am_run, ram_run  # noqa: F401
from .stats import cov, inse_mc_cov, multi_ess, acf, is_pos_def  # noqa: F401
from .philox import philox4x32_10, chain_uniforms, chain_normals  # noqa: F401
