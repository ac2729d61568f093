"""Geometry constants and vehicle-mode tables of the crossroad scenario.

Mirror of the parts of the reference's endtoend_env_utils.py the model hot path reads
(reference endtoend_env_utils.py:14-46, :73-104, :232-237); the SUMO coordinate glue of
that module is out of scope.
"""
import dataclasses
from collections import OrderedDict

L, W = 4.8, 2.0
LANE_WIDTH = 3.75
LANE_NUMBER = 3
CROSSROAD_SIZE = 50
EXPECTED_V = 8.


@dataclasses.dataclass(frozen=True)
class CrossroadConfig(object):
    """What the reference bakes in as module constants and literals, as one frozen record: geometry
    (reference endtoend_env_utils.py:14-18), EXPECTED_V, and the reward weights of compute_rewards
    (reference dynamics_and_models.py:297-298).  Defaults are the reference's values.  `set_config`
    pushes a record into the kernels' constant block (ce2e_config_set) and makes it the one new
    ReferencePath tables are built from; existing ReferencePath objects keep their tables."""
    L: float = L
    W: float = W
    LANE_WIDTH: float = LANE_WIDTH
    LANE_NUMBER: int = LANE_NUMBER
    CROSSROAD_SIZE: float = CROSSROAD_SIZE
    EXPECTED_V: float = EXPECTED_V
    w_devi_v: float = 0.05
    w_devi_y: float = 0.8
    w_devi_phi: float = 30.
    w_punish_yaw_rate: float = 0.02
    w_punish_steer: float = 5.
    w_punish_a_x: float = 0.05


_ACTIVE = CrossroadConfig()


def get_config():
    return _ACTIVE


def set_config(cfg=None):
    """Activate `cfg` (None: the reference's defaults) for all later kernel launches and path tables.
    Returns the previous record.  Process wide; do not call while kernels are in flight."""
    global _ACTIVE
    from . import _lib
    cfg = CrossroadConfig() if cfg is None else cfg
    c = _lib.Config(cfg.L, cfg.W, cfg.LANE_WIDTH, int(cfg.LANE_NUMBER), cfg.CROSSROAD_SIZE, cfg.EXPECTED_V, cfg.w_devi_v,
                    cfg.w_devi_y, cfg.w_devi_phi, cfg.w_punish_yaw_rate, cfg.w_punish_steer, cfg.w_punish_a_x)
    import ctypes
    _lib.check(_lib.load().ce2e_config_set(ctypes.byref(c)))
    old, _ACTIVE = _ACTIVE, cfg
    return old

# how many vehicles of each route class an observation holds, per ego task (EU:21-23)
VEHICLE_MODE_DICT = dict(left=OrderedDict(dl=2, du=2, ud=2, ul=2),
                         straight=OrderedDict(dl=1, du=2, ud=2, ru=2, ur=2),
                         right=OrderedDict(dr=1, ur=2, lr=2))


def dict2flat(inp):
    return [key for key, val in inp.items() for _ in range(val)]


def dict2num(inp):
    return sum(inp.values())


VEH_NUM = {task: dict2num(d) for task, d in VEHICLE_MODE_DICT.items()}
VEHICLE_MODE_LIST = {task: dict2flat(d) for task, d in VEHICLE_MODE_DICT.items()}

TASKS = ('left', 'straight', 'right')

# route classes whose vehicles turn on an arc inside the junction box
# (reference dynamics_and_models.py:416 / :418)
LEFT_TURN_MODES = frozenset(('dl', 'rd', 'ur', 'lu'))
RIGHT_TURN_MODES = frozenset(('dr', 'ru', 'ul', 'ld'))


def turn_class(mode):
    """+1 / -1 / 0: which heading-rate branch predict_for_a_mode takes for a route class."""
    if isinstance(mode, (int,)) or hasattr(mode, '__index__'):
        return int(mode)
    if mode in LEFT_TURN_MODES:
        return 1
    if mode in RIGHT_TURN_MODES:
        return -1
    return 0


def judge_feasible(orig_x, orig_y, task):
    """Is the point on the drivable area of `task` (reference endtoend_env_utils.py:73-104)."""
    half = CROSSROAD_SIZE / 2
    road = LANE_WIDTH * LANE_NUMBER
    if -half < orig_x < half and -half < orig_y < half:
        return True
    if task == 'left':
        return bool((0 < orig_x < LANE_WIDTH and orig_y <= -half) or (0 < orig_y < road and orig_x < -half))
    if task == 'straight':
        return bool((LANE_WIDTH < orig_x < 2 * LANE_WIDTH and orig_y <= -half) or
                    (0 < orig_x < road and orig_y >= half))
    assert task == 'right'
    return bool((2 * LANE_WIDTH < orig_x < road and orig_y <= -half) or (-road < orig_y < 0 and orig_x > half))


def deal_with_phi(phi):
    """Wrap a heading in degrees to (-180, 180] (reference endtoend_env_utils.py:232-237)."""
    while phi > 180:
        phi -= 360
    while phi <= -180:
        phi += 360
    return phi
