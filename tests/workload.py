"""test-side alias of the synthetic workload builders (inputs only)."""
from adsurf_b200.workload import config5, config5_ensemble, config5_periods, config5_velocities  # noqa: F401
