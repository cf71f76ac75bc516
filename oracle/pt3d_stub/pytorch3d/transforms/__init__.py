"""Stand-in leaf (TEST INFRASTRUCTURE ONLY): delegates to the oracle's restatement -- anchored against scipy in
tests/test_cpu_oracle_and_host.py::test_leaf_so3_exp_map_matches_scipy, not against pytorch3d."""
from oracle.render_oracle import so3_exp_map  # noqa: F401
