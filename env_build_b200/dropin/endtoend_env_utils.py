"""Flat-module shim for `from endtoend_env_utils import ...` (reference dynamics_and_models.py:19-20)."""
from env_build_b200.endtoend_env_utils import *  # noqa: F401,F403
from env_build_b200.endtoend_env_utils import (CROSSROAD_SIZE, EXPECTED_V, L, LANE_NUMBER, LANE_WIDTH,  # noqa: F401
                                               VEH_NUM, VEHICLE_MODE_DICT, VEHICLE_MODE_LIST, W)
