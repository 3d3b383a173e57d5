"""CPU oracle for the discrete-domain GP hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, line by line in NumPy fp64, the arithmetic of the
reference (jonhue/transductive-active-learning) for the path named in
BASELINE.json `north_star`.  It is the checker the CUDA path is compared
against; it is never on the product path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
leg may import it.

Pinning status (see DESIGN.md "Oracle"):
  * PINNED against every known-answer value the reference's own tests hold for
    this path: NLL_0 = 1.552 (RBF) and 4.560 (Laplace) of
    tests/test_hyperopt.py:29,45 (these exercise Gaussian/Laplace.__call__,
    noise_covariance_matrix, solve_linear_system, log_prob) and the exact
    prior-state checks of tests/test_discrete_model.py:29-39 and
    tests/test_continuous_model.py:30-40.
  * PARITY UNPINNED for MarginalModel.step / posterior, ITL.F, VTL.F, the
    safe-set masks and objective_argmax: the reference's tests hold no vectors
    for them and the reference itself cannot be imported in this image (jax,
    jaxlib, chex, optax, matplotlib absent; no network), so no fixtures could
    be generated from it.  Those functions are restated from the source and
    cross-checked only through mathematical identities (tests/test_oracle.py).
"""
from .ref_numpy import *  # noqa: F401,F403
