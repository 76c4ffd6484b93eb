"""CPU oracle package -- TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this package, and only as the checker or the
reported CPU baseline.  Nothing under `sim_a_splat_b200/` imports it.
"""
from .oracle import OracleFrame, build_oracle, compute_cov, exp_neg, num_threads, render, set_num_threads  # noqa: F401
from . import pose_oracle  # noqa: F401
