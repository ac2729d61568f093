"""Flat-module shim: put env_build_b200/dropin on sys.path ahead of the reference checkout and
`from dynamics_and_models import VehicleDynamics, ReferencePath, EnvironmentModel` (reference
endtoend.py:20, hier_decision.py:21, multi_ego.py:20, mpc_ipopt.py:16) resolves to the B200 path."""
from env_build_b200.dynamics_and_models import (DeviceTensor, EnvironmentModel, ReferencePath,  # noqa: F401
                                                VehicleDynamics, deal_with_phi_diff)
