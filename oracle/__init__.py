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
  * PINNED against outputs of the reference's OWN SOURCE run in the build
    container: /root/reference/lib is imported and executed over a NumPy
    stand-in for jax.numpy / jax.random / chex (tests/golden/refshim.py; jax and
    jaxlib 0.4.30 are absent and not installable).  The committed vectors
    (tests/golden/reference_calls.npz, reference_trajectories.npz, generator
    tests/golden/make_reference_fixtures.py) cover the kernels, get_indices /
    get_mask / is_subset_of, compute_posterior_covariance / posterior,
    MarginalModel.step (m = 0, 1, 5, heteroscedastic noise, callable beta), all
    safe-set masks, sqd_correlation, objective_argmax (messages, tie set) and
    eight BO trajectories (ITL, VTL, CTL, MM-ITL with PM / fixed ROIs);
    tests/test_reference_fixtures.py holds the oracle to them at 1e-10.
    What those vectors do NOT pin is XLA's own arithmetic (fusion, Eigen's
    summation order, threefry draws): rounding-level differences and the choice
    among exact ties (jr.choice) remain documented, not reproduced.
"""
from .ref_numpy import *  # noqa: F401,F403
