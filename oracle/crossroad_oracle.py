"""CPU oracle: NumPy fp32 restatement of the reference CrossroadEnd2end model hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package.
The product (env_build_b200/) never does; it fails loudly without its CUDA
extension.

What it restates (all citations into /root/reference/):
  dynamics_and_models.py:26-87    VehicleDynamics.__init__/f_xu/prediction
  dynamics_and_models.py:118-132  EnvironmentModel.rollout_out / action scaling
  dynamics_and_models.py:134-184  EnvironmentModel.ss
  dynamics_and_models.py:186-320  EnvironmentModel.compute_rewards
  dynamics_and_models.py:322-358  EnvironmentModel.compute_next_obses
  dynamics_and_models.py:386-427  ego_predict / veh_predict / predict_for_a_mode
  dynamics_and_models.py:577-770  deal_with_phi_diff, ReferencePath
  endtoend_env_utils.py:14-46     constants and vehicle-mode tables
  endtoend_env_utils.py:73-104    judge_feasible
  endtoend.py:132-283, 501-507    Gym step: action scaling, next ego state, ego dynamics (corner
                                  points, r_bound), done logic, reward (gym_* functions)
  traffic.py:263-295              two-circle collision check
  endtoend.py:340-464             interested-vehicle selection (select_interested_vehicles)

Numerics policy: tensors are fp32; + - * / sqrt are the
IEEE fp32 ops NumPy performs (one rounding each, no FMA), in the association
order of the reference source; Python scalars are rounded to fp32 where TF
would do so (when first combined with a tensor); sin / cos / atan / atan2 are
evaluated in float64 and rounded to fp32 (TF-Eigen's fp32 kernels are ~1 ulp
approximations of exactly that value and cannot be reproduced offline);
argmin returns the first minimum.

Third-party arithmetic absent from /root/reference: the `bezier` PyPI package
(un-pinned; no requirements file).  Its published cubic evaluation
(`evaluate_multi_barycentric`, float64) is restated in `_bezier_cubic`.

A float64 PyTorch restatement of rollout_out for GRADIENT checks lives in oracle/torch_model.py.

PARITY PIN: the reference holds no golden vectors or numeric asserts for this
path, and TensorFlow cannot be installed here.  The
oracle is pinned instead against the UNMODIFIED reference source executed on
a NumPy-backed TensorFlow stand-in (tests/golden/make_golden.py ->
tests/golden/*.npz; tests/test_oracle_golden.py requires bit-equality); the Gym-side
functions (next ego state, ego dynamics, done logic, reward, vehicle selection) are
pinned against the unmodified CrossroadEnd2end / Traffic METHODS run on import-only
stand-ins for gym / traci / sumolib (tests/golden/make_golden_env.py -> env_*.npz).  With
respect to a real TF2 run the parity is therefore "unpinned" at the level of
TF's transcendental kernels (<= ~1 ulp of fp32).
"""
import numpy as np

f32 = np.float32

# ----------------------------------------------------------------------------
# constants -- endtoend_env_utils.py:14-46
# ----------------------------------------------------------------------------
L, W = 4.8, 2.0
LANE_WIDTH = 3.75
LANE_NUMBER = 3
CROSSROAD_SIZE = 50
EXPECTED_V = 8.

VEHICLE_MODE_LIST = dict(
    left=['dl', 'dl', 'du', 'du', 'ud', 'ud', 'ul', 'ul'],
    straight=['dl', 'du', 'du', 'ud', 'ud', 'ru', 'ru', 'ur', 'ur'],
    right=['dr', 'ur', 'ur', 'lr', 'lr'])
VEH_NUM = {k: len(v) for k, v in VEHICLE_MODE_LIST.items()}

LEFT_TURN_MODES = ('dl', 'rd', 'ur', 'lu')    # dynamics_and_models.py:416
RIGHT_TURN_MODES = ('dr', 'ru', 'ul', 'ld')   # dynamics_and_models.py:418

TASKS = ('left', 'straight', 'right')


def turn_class(mode):
    """+1 / -1 / 0 for the three branches of predict_for_a_mode (DM:416-421)."""
    if mode in LEFT_TURN_MODES:
        return 1
    if mode in RIGHT_TURN_MODES:
        return -1
    return 0


def _sin(x):
    return np.sin(np.asarray(x, dtype=np.float64)).astype(f32)


def _cos(x):
    return np.cos(np.asarray(x, dtype=np.float64)).astype(f32)


def _atan(x):
    return np.arctan(np.asarray(x, dtype=np.float64)).astype(f32)


def _sq(x):
    return x * x


# ----------------------------------------------------------------------------
# VehicleDynamics -- dynamics_and_models.py:26-87
# ----------------------------------------------------------------------------
VEHICLE_PARAMS = dict(C_f=-155495.0, C_r=-155495.0, a=1.19, b=1.46, mass=1520., I_z=2642., miu=0.8, g=9.81)
VEHICLE_PARAMS['F_zf'] = (VEHICLE_PARAMS['b'] * VEHICLE_PARAMS['mass'] * VEHICLE_PARAMS['g'] /
                          (VEHICLE_PARAMS['a'] + VEHICLE_PARAMS['b']))   # float64, DM:48
VEHICLE_PARAMS['F_zr'] = (VEHICLE_PARAMS['a'] * VEHICLE_PARAMS['mass'] * VEHICLE_PARAMS['g'] /
                          (VEHICLE_PARAMS['a'] + VEHICLE_PARAMS['b']))


def f_xu(states, actions, tau):
    """DM:52-83.  states [B,6] (phi in degrees), actions [B,2] scaled (steer rad, a_x), tau Python float.
    Returns (next_states [B,6], params [B,4] = alpha_f, alpha_r, miu_f, miu_r)."""
    states = np.asarray(states, dtype=f32)
    actions = np.asarray(actions, dtype=f32)
    v_x, v_y, r, x, y, phi = (states[:, i] for i in range(6))
    phi = phi * f32(np.pi) / f32(180.)                          # DM:54
    steer, a_x = actions[:, 0], actions[:, 1]
    C_f, C_r = f32(VEHICLE_PARAMS['C_f']), f32(VEHICLE_PARAMS['C_r'])
    a, b = f32(VEHICLE_PARAMS['a']), f32(VEHICLE_PARAMS['b'])
    mass, I_z = f32(VEHICLE_PARAMS['mass']), f32(VEHICLE_PARAMS['I_z'])
    miu, g = f32(VEHICLE_PARAMS['miu']), f32(VEHICLE_PARAMS['g'])
    tau = f32(tau)

    F_zf, F_zr = b * mass * g / (a + b), a * mass * g / (a + b)                     # DM:65
    zeros = np.zeros_like(a_x)
    F_xf = np.where(a_x < 0, mass * a_x / f32(2), zeros)                            # DM:66
    F_xr = np.where(a_x < 0, mass * a_x / f32(2), mass * a_x)                       # DM:67
    with np.errstate(invalid='ignore', divide='ignore'):
        miu_f = np.sqrt(_sq(miu * F_zf) - _sq(F_xf)) / F_zf                         # DM:68
        miu_r = np.sqrt(_sq(miu * F_zr) - _sq(F_xr)) / F_zr                         # DM:69
        alpha_f = _atan((v_y + a * r) / (v_x + f32(1e-8))) - steer                  # DM:70
        alpha_r = _atan((v_y - b * r) / (v_x + f32(1e-8)))                          # DM:71

        cphi, sphi = _cos(phi), _sin(phi)
        K1 = a * C_f - b * C_r
        nxt = [v_x + tau * (a_x + v_y * r),                                          # DM:73
               (mass * v_y * v_x + tau * K1 * r - tau * C_f * steer * v_x
                - tau * mass * _sq(v_x) * r) / (mass * v_x - tau * (C_f + C_r)),    # DM:74-76
               (-I_z * r * v_x - tau * K1 * v_y + tau * a * C_f * steer * v_x) /
               (tau * (_sq(a) * C_f + _sq(b) * C_r) - I_z * v_x),                   # DM:77-78
               x + tau * (v_x * cphi - v_y * sphi),                                 # DM:79
               y + tau * (v_x * sphi + v_y * cphi),                                 # DM:80
               (phi + tau * r) * f32(180) / f32(np.pi)]                             # DM:81
    return np.stack(nxt, 1).astype(f32), np.stack([alpha_f, alpha_r, miu_f, miu_r], 1).astype(f32)


def prediction(x_1, u_1, frequency):
    """DM:85-87."""
    return f_xu(x_1, u_1, 1 / frequency)


def action_transformation(actions):
    """DM:128-132 (NumPy twin endtoend.py:258-267): clip +-1.05; steer=0.4*a0; a_x=2.25*a1-0.75."""
    actions = np.asarray(actions, dtype=f32)
    actions = np.minimum(np.maximum(actions, f32(-1.05)), f32(1.05))
    steer_norm, a_xs_norm = actions[:, 0], actions[:, 1]
    return np.stack([f32(0.4) * steer_norm, f32(2.25) * a_xs_norm - f32(0.75)], 1)


# ----------------------------------------------------------------------------
# ReferencePath -- dynamics_and_models.py:577-770
# ----------------------------------------------------------------------------
def deal_with_phi_diff(phi_diff):
    """DM:577-580: one wrap each side, not a modulo."""
    phi_diff = np.where(phi_diff > f32(180.), phi_diff - f32(360.), phi_diff)
    phi_diff = np.where(phi_diff < f32(-180.), phi_diff + f32(360.), phi_diff)
    return phi_diff


def _bezier_cubic(nodes32, s_vals):
    """`bezier.Curve(nodes, degree=3).evaluate_multi(s_vals)` (DM:616-618): the
    package's barycentric Horner scheme, float64.  nodes32: [2,4] float32."""
    nodes = np.asarray(nodes32, dtype=np.float64)
    s = np.asarray(s_vals, dtype=np.float64)
    l1 = (1.0 - s)[np.newaxis, :]
    l2 = s[np.newaxis, :]
    result = np.zeros((2, s.shape[0]))
    result += l1 * nodes[:, [0]]
    binom = 1.0
    l2_pow = np.ones((1, s.shape[0]))
    for index in range(1, 3):
        l2_pow = l2_pow * l2
        binom = (binom * (3 - index + 1)) / index
        result += binom * l2_pow * nodes[:, [index]]
        result *= l1
    result += l2 * l2_pow * nodes[:, [3]]
    return result


def construct_ref_paths(task):
    """DM:598-700.  Returns (path_list, path_len_list, control_points); each path
    is a tuple (x, y, phi_deg) of equal-length float32 arrays.

    Heading: the reference calls np.arctan2 on float32 arrays (DM:629), whose
    last-ulp result depends on the NumPy build's SIMD dispatch.  The oracle
    uses float64 atan2 rounded to fp32 (machine independent); see
    tests/test_oracle_golden.py for the measured ulp distance."""
    from math import pi
    sl = 40
    ratio = 30
    half = CROSSROAD_SIZE / 2
    control_ext = CROSSROAD_SIZE / 3.
    path_list, len_list, ctrl = [], [], []
    assert task in TASKS

    if task == 'left':
        end_offsets = [LANE_WIDTH * (i + 0.5) for i in range(LANE_NUMBER)]
        start_offset = LANE_WIDTH * 0.5
        n_curve = int(pi / 2 * (half + LANE_WIDTH / 2)) * ratio
    elif task == 'straight':
        end_offsets = [LANE_WIDTH * (i + 0.5) for i in range(LANE_NUMBER)]
        start_offset = LANE_WIDTH * 1.5
        n_curve = CROSSROAD_SIZE * ratio
    else:
        control_ext = CROSSROAD_SIZE / 5.
        end_offsets = [-LANE_WIDTH * 2.5, -LANE_WIDTH * 1.5, -LANE_WIDTH * 0.5]
        start_offset = LANE_WIDTH * (LANE_NUMBER - 0.5)
        n_curve = int(pi / 2 * (half - LANE_WIDTH * (LANE_NUMBER - 0.5))) * ratio

    n_line = sl * ratio
    for end_offset in end_offsets:
        cp1 = (start_offset, -half)
        cp2 = (start_offset, -half + control_ext)
        if task == 'left':
            cp3, cp4 = (-half + control_ext, end_offset), (-half, end_offset)
        elif task == 'straight':
            cp3, cp4 = (end_offset, half - control_ext), (end_offset, half)
        else:
            cp3, cp4 = (half - control_ext, end_offset), (half, end_offset)
        ctrl.append([cp1, cp2, cp3, cp4])
        node = np.array([[cp1[0], cp2[0], cp3[0], cp4[0]],
                         [cp1[1], cp2[1], cp3[1], cp4[1]]], dtype=f32)
        trj = _bezier_cubic(node, np.linspace(0, 1.0, n_curve)).astype(f32)

        start_x = (start_offset * np.ones(shape=(n_line,), dtype=f32))[:-1]       # DM:620 / 653 / 688
        start_y = np.linspace(-half - sl, -half, n_line, dtype=f32)[:-1]
        if task == 'left':
            end_x = np.linspace(-half, -half - sl, n_line, dtype=f32)[1:]
            end_y = (end_offset * np.ones(shape=(n_line,), dtype=f32))[1:]
        elif task == 'straight':
            end_x = (end_offset * np.ones(shape=(n_line,), dtype=f32))[1:]
            end_y = np.linspace(half, half + sl, n_line, dtype=f32)[1:]
        else:
            end_x = np.linspace(half, half + sl, n_line, dtype=f32)[1:]
            end_y = (end_offset * np.ones(shape=(n_line,), dtype=f32))[1:]
        xs = np.append(np.append(start_x, trj[0]), end_x)
        ys = np.append(np.append(start_y, trj[1]), end_y)
        xs_1, ys_1 = xs[:-1], ys[:-1]
        xs_2, ys_2 = xs[1:], ys[1:]
        dy, dx = ys_2 - ys_1, xs_2 - xs_1                                          # fp32 differences
        ang = np.arctan2(dy.astype(np.float64), dx.astype(np.float64)).astype(f32)
        phis_1 = ang * f32(180) / f32(pi)                                          # DM:629-630
        path_list.append((xs_1, ys_1, phis_1))
        len_list.append((n_line, n_curve, len(xs_1)))
    return path_list, len_list, ctrl


class ReferencePath(object):
    """DM:583-770 on NumPy arrays.  `path_list` may be injected (used by the
    golden tests to take the tables produced by the reference itself)."""

    def __init__(self, task, ref_index=None, path_list=None):
        self.exp_v = EXPECTED_V
        self.task = task
        if path_list is None:
            self.path_list, self.path_len_list, self.control_points = construct_ref_paths(task)
        else:
            self.path_list = [tuple(np.asarray(a, dtype=f32) for a in p) for p in path_list]
            self.path_len_list, self.control_points = None, None
        self.ref_index = np.random.choice(len(self.path_list)) if ref_index is None else ref_index
        self.path = self.path_list[self.ref_index]

    def set_path(self, path_index=None):
        self.ref_index = path_index
        self.path = self.path_list[self.ref_index]

    def find_closest_point(self, xs, ys, ratio=10, return_margin=False):
        """DM:702-715: decimate by `ratio`, brute-force squared distance, FIRST argmin."""
        xs, ys = np.asarray(xs, dtype=f32), np.asarray(ys, dtype=f32)
        path_len = len(self.path[0])
        reduced_idx = np.arange(0, path_len, ratio)
        rx, ry = self.path[0][reduced_idx], self.path[1][reduced_idx]
        dist = _sq(xs[:, None] - rx[None, :]) + _sq(ys[:, None] - ry[None, :])      # DM:712
        amin = np.argmin(dist, 1)
        indexs = amin.astype(np.int64) * ratio
        if return_margin:
            # test aid: gap between the best and the second-best candidate (near-tie detector)
            part = np.partition(dist, 1, axis=1) if dist.shape[1] > 1 else np.concatenate([dist, dist + 1], 1)
            margin = part[:, 1] - part[:, 0]
            return indexs, self.indexs2points(indexs), margin
        return indexs, self.indexs2points(indexs)

    def future_n_data(self, current_indexs, n):
        """DM:717-724."""
        out = []
        current_indexs = np.asarray(current_indexs).astype(np.int32)
        for _ in range(n):
            current_indexs = current_indexs + 80
            current_indexs = np.where(current_indexs >= len(self.path[0]) - 2, len(self.path[0]) - 2,
                                      current_indexs).astype(np.int32)
            out.append(self.indexs2points(current_indexs))
        return out

    def indexs2points(self, indexs):
        """DM:726-733."""
        indexs = np.asarray(indexs)
        indexs = np.where(indexs >= 0, indexs, 0)
        indexs = np.where(indexs < len(self.path[0]), indexs, len(self.path[0]) - 1)
        return self.path[0][indexs], self.path[1][indexs], self.path[2][indexs]

    def _two2one(self, ego_xs, ego_ys, ref_xs, ref_ys):
        """DM:736-752 (returns -delta)."""
        half = f32(CROSSROAD_SIZE / 2)
        if self.task == 'left':
            delta_ = np.sqrt(_sq(ego_xs - (-half)) + _sq(ego_ys - (-half))) - \
                np.sqrt(_sq(ref_xs - (-half)) + _sq(ref_ys - (-half)))
            delta_ = np.where(ego_ys < -half, ego_xs - ref_xs, delta_)
            delta_ = np.where(ego_xs < -half, ego_ys - ref_ys, delta_)
            return -delta_
        elif self.task == 'straight':
            return -(ego_xs - ref_xs)
        else:
            assert self.task == 'right'
            delta_ = -(np.sqrt(_sq(ego_xs - half) + _sq(ego_ys - (-half))) -
                       np.sqrt(_sq(ref_xs - half) + _sq(ref_ys - (-half))))
            delta_ = np.where(ego_ys < -half, ego_xs - ref_xs, delta_)
            delta_ = np.where(ego_xs > half, -(ego_ys - ref_ys), delta_)
            return -delta_

    def tracking_error_vector(self, ego_xs, ego_ys, ego_phis, ego_vs, n, return_margin=False):
        """DM:735-770 -> [B, 3(n+1)]."""
        ego_xs, ego_ys = np.asarray(ego_xs, dtype=f32), np.asarray(ego_ys, dtype=f32)
        ego_phis, ego_vs = np.asarray(ego_phis, dtype=f32), np.asarray(ego_vs, dtype=f32)
        res = self.find_closest_point(ego_xs, ego_ys, return_margin=return_margin)
        indexs, current_points = res[0], res[1]
        n_future_data = self.future_n_data(indexs, n)
        tracking_error = np.stack([self._two2one(ego_xs, ego_ys, current_points[0], current_points[1]),
                                   deal_with_phi_diff(ego_phis - current_points[2]),
                                   ego_vs - f32(self.exp_v)], 1)
        final = tracking_error
        if n > 0:
            future_points = np.concatenate([np.stack([ref_point[0] - ego_xs,
                                                      ref_point[1] - ego_ys,
                                                      deal_with_phi_diff(ego_phis - ref_point[2])], 1)
                                            for ref_point in n_future_data], 1)
            final = np.concatenate([final, future_points], 1)
        final = final.astype(f32)
        if return_margin:
            return final, res[2]
        return final


# ----------------------------------------------------------------------------
# EnvironmentModel -- dynamics_and_models.py:90-427
# ----------------------------------------------------------------------------
def _circle_points(x, y, phi_deg):
    """Front / rear circle centres at +-(L-W)/2 along the heading (DM:210-214, 220-224)."""
    lws = f32((L - W) / 2.)
    ang = phi_deg * f32(np.pi) / f32(180.)
    c, s = _cos(ang), _sin(ang)
    return (x + lws * c, y + lws * s), (x - lws * c, y - lws * s)


def _road_terms(task, point, real):
    """One ego point's four road-edge hinge terms, source order (DM:233-295)."""
    px, py = point
    half = f32(CROSSROAD_SIZE / 2)
    lw = f32(LANE_WIDTH)
    lw2 = f32(2 * LANE_WIDTH)
    lw3 = f32(LANE_WIDTH * LANE_NUMBER)
    one = f32(1)
    z = np.zeros_like(px)
    out = []
    if task == 'left':
        out.append(np.where(np.logical_and(py < -half, px < one), _sq(px - one), z))
        out.append(np.where(np.logical_and(py < -half, lw - px < one), _sq(lw - px - one), z))
        third_cond = (px < -half) if real else (px < f32(0))                    # DM:248 vs DM:239
        out.append(np.where(np.logical_and(third_cond, lw3 - py < one), _sq(lw3 - py - one), z))
        out.append(np.where(np.logical_and(px < -half, py - f32(0) < one), _sq(py - f32(0) - one), z))
    elif task == 'straight':
        out.append(np.where(np.logical_and(py < -half, px - lw < one), _sq(px - lw - one), z))
        out.append(np.where(np.logical_and(py < -half, lw2 - px < one), _sq(lw2 - px - one), z))
        out.append(np.where(np.logical_and(py > half, lw3 - px < one), _sq(lw3 - px - one), z))
        out.append(np.where(np.logical_and(py > half, px - f32(0) < one), _sq(px - f32(0) - one), z))
    else:
        assert task == 'right'
        out.append(np.where(np.logical_and(py < -half, px - lw2 < one), _sq(px - lw2 - one), z))
        out.append(np.where(np.logical_and(py < -half, lw3 - px < one), _sq(lw3 - px - one), z))
        out.append(np.where(np.logical_and(px > half, f32(0) - py < one), _sq(f32(0) - py - one), z))
        out.append(np.where(np.logical_and(px > half, py - (-lw3) < one), _sq(py - (-lw3) - one), z))
    return out


def compute_rewards(obses, actions, task, num_future_data=0):
    """DM:186-320.  `actions` are the SCALED actions.  Returns
    (rewards, punish_term_for_training, real_punish_term, veh2veh4real, veh2road4real, reward_dict)."""
    obses = np.asarray(obses, dtype=f32)
    actions = np.asarray(actions, dtype=f32)
    ntr = 3 * (num_future_data + 1)
    ego_infos, tracking_infos, veh_infos = obses[:, :6], obses[:, 6:6 + ntr], obses[:, 6 + ntr:]
    steers, a_xs = actions[:, 0], actions[:, 1]
    punish_steer = -_sq(steers)
    punish_a_x = -_sq(a_xs)
    punish_yaw_rate = -_sq(ego_infos[:, 2])
    devi_y = -_sq(tracking_infos[:, 0])
    devi_phi = -_sq(tracking_infos[:, 1] * f32(np.pi) / f32(180.))
    devi_v = -_sq(tracking_infos[:, 2])

    ego_front, ego_rear = _circle_points(ego_infos[:, 3], ego_infos[:, 4], ego_infos[:, 5])
    zeros = np.zeros_like(ego_infos[:, 0])
    veh2veh4real = zeros.copy()
    veh2veh4training = zeros.copy()
    for veh_index in range(int(veh_infos.shape[1] / 4)):                           # DM:218
        vehs = veh_infos[:, veh_index * 4:(veh_index + 1) * 4]
        veh_front, veh_rear = _circle_points(vehs[:, 0], vehs[:, 1], vehs[:, 3])
        for ego_point in (ego_front, ego_rear):
            for veh_point in (veh_front, veh_rear):
                d = np.sqrt(_sq(ego_point[0] - veh_point[0]) + _sq(ego_point[1] - veh_point[1]))
                veh2veh4training = veh2veh4training + np.where(d - f32(3.5) < 0, _sq(d - f32(3.5)), zeros)
                veh2veh4real = veh2veh4real + np.where(d - f32(2.5) < 0, _sq(d - f32(2.5)), zeros)

    veh2road4real = zeros.copy()
    veh2road4training = zeros.copy()
    for ego_point in (ego_front, ego_rear):
        for t in _road_terms(task, ego_point, real=False):
            veh2road4training = veh2road4training + t
        for t in _road_terms(task, ego_point, real=True):
            veh2road4real = veh2road4real + t

    rewards = f32(0.05) * devi_v + f32(0.8) * devi_y + f32(30) * devi_phi + f32(0.02) * punish_yaw_rate + \
        f32(5) * punish_steer + f32(0.05) * punish_a_x                              # DM:297-298
    punish_term_for_training = veh2veh4training + veh2road4training
    real_punish_term = veh2veh4real + veh2road4real
    reward_dict = dict(punish_steer=punish_steer, punish_a_x=punish_a_x, punish_yaw_rate=punish_yaw_rate,
                       devi_v=devi_v, devi_y=devi_y, devi_phi=devi_phi,
                       scaled_punish_steer=f32(5) * punish_steer,
                       scaled_punish_a_x=f32(0.05) * punish_a_x,
                       scaled_punish_yaw_rate=f32(0.02) * punish_yaw_rate,
                       scaled_devi_v=f32(0.05) * devi_v,
                       scaled_devi_y=f32(0.8) * devi_y,
                       scaled_devi_phi=f32(30) * devi_phi,
                       veh2veh4training=veh2veh4training, veh2road4training=veh2road4training,
                       veh2veh4real=veh2veh4real, veh2road4real=veh2road4real)
    return rewards, punish_term_for_training, real_punish_term, veh2veh4real, veh2road4real, reward_dict


REWARD_DICT_KEYS = ('punish_steer', 'punish_a_x', 'punish_yaw_rate', 'devi_v', 'devi_y', 'devi_phi',
                    'scaled_punish_steer', 'scaled_punish_a_x', 'scaled_punish_yaw_rate', 'scaled_devi_v',
                    'scaled_devi_y', 'scaled_devi_phi', 'veh2veh4training', 'veh2road4training',
                    'veh2veh4real', 'veh2road4real')


def predict_for_a_mode(vehs, mode):
    """DM:405-427.  `mode` is a route string ('dl', ...) or a turn class (+1/-1/0)."""
    tc = mode if isinstance(mode, (int, np.integer)) else turn_class(mode)
    vehs = np.asarray(vehs, dtype=f32)
    veh_xs, veh_ys, veh_vs, veh_phis = vehs[:, 0], vehs[:, 1], vehs[:, 2], vehs[:, 3]
    veh_phis_rad = veh_phis * f32(np.pi) / f32(180.)
    half = f32(CROSSROAD_SIZE / 2)
    middle_cond = np.logical_and(np.logical_and(veh_xs > -half, veh_xs < half),
                                 np.logical_and(veh_ys > -half, veh_ys < half))
    zeros = np.zeros_like(veh_xs)
    freq = f32(10.)
    veh_xs_delta = veh_vs / freq * _cos(veh_phis_rad)
    veh_ys_delta = veh_vs / freq * _sin(veh_phis_rad)
    if tc > 0:
        delta = np.where(middle_cond, (veh_vs / f32(CROSSROAD_SIZE / 2 + 0.5 * LANE_WIDTH)) / freq, zeros)
    elif tc < 0:
        delta = np.where(middle_cond, -(veh_vs / f32(CROSSROAD_SIZE / 2 - 2.5 * LANE_WIDTH)) / freq, zeros)
    else:
        delta = zeros
    nx, ny, nv, nphi = veh_xs + veh_xs_delta, veh_ys + veh_ys_delta, veh_vs, veh_phis_rad + delta
    nphi = np.where(nphi > f32(np.pi), nphi - f32(2 * np.pi), nphi)
    nphi = np.where(nphi <= f32(-np.pi), nphi + f32(2 * np.pi), nphi)
    nphi_deg = nphi * f32(180) / f32(np.pi)
    return np.stack([nx, ny, nv, nphi_deg], 1).astype(f32)


def veh_predict(veh_infos, mode_list):
    """DM:394-403: only the first len(mode_list) vehicles are predicted (and returned)."""
    veh_infos = np.asarray(veh_infos, dtype=f32)
    out = [predict_for_a_mode(veh_infos[:, i * 4:(i + 1) * 4], mode_list[i]) for i in range(len(mode_list))]
    return np.concatenate(out, 1) if out else np.zeros((len(veh_infos), 0), f32)


def ego_predict(ego_infos, actions):
    """DM:386-392."""
    nxt, _ = prediction(np.asarray(ego_infos, dtype=f32)[:, :6], actions, 10.)
    nxt = nxt.copy()
    nxt[:, 0] = np.minimum(np.maximum(nxt[:, 0], f32(0.)), f32(35.))
    return nxt


class EnvironmentModel(object):
    """DM:90-427 (render excluded).  `veh_mode_list` overrides VEHICLE_MODE_LIST[task]
    (needed for the V=32 synthetic configuration)."""

    def __init__(self, training_task, num_future_data=0, mode='training', veh_mode_list=None, path_list=None):
        self.task = training_task
        self.mode = mode
        self.base_frequency = 10.
        self.obses = None
        self.actions = None
        self.ref_path = ReferencePath(self.task, ref_index=0, path_list=path_list)
        self.ref_indexes = None
        self.num_future_data = num_future_data
        self.exp_v = EXPECTED_V
        self.ego_info_dim = 6
        self.per_veh_info_dim = 4
        self.per_tracking_info_dim = 3
        self.veh_mode_list = list(VEHICLE_MODE_LIST[self.task]) if veh_mode_list is None else list(veh_mode_list)

    def reset(self, obses, ref_indexes=None):
        self.obses = np.asarray(obses, dtype=f32)
        self.ref_indexes = ref_indexes
        self.actions = None

    def add_traj(self, obses, path_index):
        self.obses = np.asarray(obses, dtype=f32)
        self.ref_path.set_path(path_index)

    def rollout_out(self, actions):
        self.actions = action_transformation(actions)
        rewards, p_train, p_real, v2v_real, v2r_real, _ = self.compute_rewards(self.obses, self.actions)
        self.obses = self.compute_next_obses(self.obses, self.actions)
        return self.obses, rewards, p_train, p_real, v2v_real, v2r_real

    def compute_rewards(self, obses, actions):
        return compute_rewards(obses, actions, self.task, self.num_future_data)

    def compute_next_obses(self, obses, actions, return_margin=False):
        """DM:322-358."""
        obses = np.asarray(obses, dtype=f32)
        ntr = 3 * (self.num_future_data + 1)
        ego_infos, veh_infos = obses[:, :6], obses[:, 6 + ntr:]
        next_ego = ego_predict(ego_infos, actions)
        margins = None
        if self.mode != 'training':
            res = self.ref_path.tracking_error_vector(next_ego[:, 3], next_ego[:, 4], next_ego[:, 5],
                                                      next_ego[:, 0], self.num_future_data,
                                                      return_margin=return_margin)
            next_tracking, margins = res if return_margin else (res, None)
        else:
            next_tracking = np.zeros((len(next_ego), ntr), dtype=f32)
            margins = np.full((len(next_ego),), np.inf, dtype=f32)
            ref_indexes = np.asarray(self.ref_indexes)[:, None]
            for ref_idx, path in enumerate(self.ref_path.path_list):
                self.ref_path.path = path
                res = self.ref_path.tracking_error_vector(next_ego[:, 3], next_ego[:, 4], next_ego[:, 5],
                                                          next_ego[:, 0], self.num_future_data,
                                                          return_margin=return_margin)
                t, m = res if return_margin else (res, None)
                next_tracking = np.where(ref_indexes == ref_idx, t, next_tracking)
                if return_margin:
                    margins = np.where(ref_indexes[:, 0] == ref_idx, m, margins)
        next_veh = veh_predict(veh_infos, self.veh_mode_list)
        out = np.concatenate([next_ego, next_tracking, next_veh], 1).astype(f32)
        if return_margin:
            return out, margins
        return out

    def ss(self, obses, actions, lam=0.1):
        """DM:134-184 discrete barrier penalty."""
        obses = np.asarray(obses, dtype=f32)
        actions = action_transformation(actions)
        next_obses = self.compute_next_obses(obses, actions)
        ntr = 3 * (self.num_future_data + 1)
        ego, veh = obses[:, :6], obses[:, 6 + ntr:]
        nego, nveh = next_obses[:, :6], next_obses[:, 6 + ntr:]
        ef, er = _circle_points(ego[:, 3], ego[:, 4], ego[:, 5])
        nef, ner = _circle_points(nego[:, 3], nego[:, 4], nego[:, 5])
        out = np.zeros_like(ego[:, 0])
        one_m_lam = f32(1 - lam)
        for vi in range(int(veh.shape[1] / 4)):
            vehs = veh[:, vi * 4:(vi + 1) * 4]
            ego2veh = np.sqrt(_sq(ego[:, 3] - vehs[:, 0]) + _sq(ego[:, 4] - vehs[:, 1]))
            nvehs = nveh[:, vi * 4:(vi + 1) * 4]
            vf, vr = _circle_points(vehs[:, 0], vehs[:, 1], vehs[:, 3])
            nvf, nvr = _circle_points(nvehs[:, 0], nvehs[:, 1], nvehs[:, 3])
            for ep in ((ef, nef), (er, ner)):
                for vp in ((vf, nvf), (vr, nvr)):
                    d = np.sqrt(_sq(ep[0][0] - vp[0][0]) + _sq(ep[0][1] - vp[0][1]))
                    nd = np.sqrt(_sq(ep[1][0] - vp[1][0]) + _sq(ep[1][1] - vp[1][1]))
                    next_g = nd - f32(2.5)
                    g = d - f32(2.5)
                    h = next_g - one_m_lam * g
                    out = out + np.where(np.logical_and(h < 0, ego2veh < f32(10)), _sq(h), np.zeros_like(out))
        return out


# ----------------------------------------------------------------------------
# Gym-side numeric helpers -- endtoend.py / endtoend_env_utils.py
# ----------------------------------------------------------------------------
def deal_with_phi(phi):
    """endtoend_env_utils.py:232-237 (scalar, float64 like the reference)."""
    while phi > 180:
        phi -= 360
    while phi <= -180:
        phi += 360
    return phi


def judge_feasible(orig_x, orig_y, task):
    """endtoend_env_utils.py:73-104 (scalar)."""
    half = CROSSROAD_SIZE / 2
    in_middle = (-half < orig_y < half) and (-half < orig_x < half)
    if task == 'left':
        before = 0 < orig_x < LANE_WIDTH and orig_y <= -half
        after = 0 < orig_y < LANE_WIDTH * LANE_NUMBER and orig_x < -half
    elif task == 'straight':
        before = LANE_WIDTH < orig_x < LANE_WIDTH * 2 and orig_y <= -half
        after = 0 < orig_x < LANE_WIDTH * LANE_NUMBER and orig_y >= half
    else:
        assert task == 'right'
        before = LANE_WIDTH * 2 < orig_x < LANE_WIDTH * 3 and orig_y <= -half
        after = -LANE_WIDTH * LANE_NUMBER < orig_y < 0 and orig_x > half
    return bool(before or after or in_middle)


def rotate_and_shift_coordination(orig_x, orig_y, orig_d, shift_x, shift_y, rotate_d):
    """endtoend_env_utils.py:120-154 on float64 arrays (angle output omitted)."""
    rad = rotate_d * np.pi / 180
    tx = orig_x * np.cos(rad) + orig_y * np.sin(rad)
    ty = -orig_x * np.sin(rad) + orig_y * np.cos(rad)
    return tx - shift_x, ty - shift_y


DONE_CODES = ('not_done_yet', 'collision', 'break_road_constrain', 'deviate_too_much', 'break_stability',
              'break_red_light', 'good_done')


def gym_next_ego_state(ego6, actions_norm):
    """CrossroadEnd2end._action_transformation_for_end2end + _get_next_ego_state (E2E:258-283):
    returns (scaled actions, next ego state with v_x floored at 0 and heading wrapped, tyre params)."""
    scaled = action_transformation(actions_norm)
    nxt, params = f_xu(np.asarray(ego6, dtype=f32)[:, :6], scaled, 1 / 10)          # E2E:279
    nxt = nxt.copy()
    nxt[:, 0] = np.where(nxt[:, 0] >= 0, nxt[:, 0], f32(0))                         # E2E:281
    nxt[:, 5] = np.array([deal_with_phi(float(p)) for p in nxt[:, 5]], dtype=f32)   # E2E:282
    return scaled, nxt, params


def gym_ego_dynamics(nxt, params):
    """CrossroadEnd2end._get_ego_dynamics (E2E:150-183): the four body corners [B,4,2] and r_bound
    [B], float64 like the reference's Python scalars (legacy NumPy promotes float32-scalar x
    Python-float to float64)."""
    vx, x, y, phi = (np.asarray(nxt)[:, i].astype(np.float64) for i in (0, 3, 4, 5))
    miu_r = np.asarray(params)[:, 3].astype(np.float64)
    r_bound = miu_r * VEHICLE_PARAMS['g'] / (np.abs(vx) + 1e-8)                      # E2E:167
    corners = []
    for lx, ly in ((L / 2, W / 2), (L / 2, -W / 2), (-L / 2, W / 2), (-L / 2, -W / 2)):   # E2E:171-177
        px, py = rotate_and_shift_coordination(lx, ly, 0, -x, -y, -phi)
        corners.append(np.stack([px, py], 1))
    return np.stack(corners, 1), r_bound


def gym_judge_done(nxt, params, delta_y, veh_after, task, v_light=0):
    """CrossroadEnd2end._judge_done (E2E:200-256) with Traffic.collision_check (traffic.py:263-295,
    every surrounding vehicle sized L x W like the ego).  veh_after [B, 4V] = vehicles after the
    traffic step; v_light scalar or [B].  Returns (code [B] int8 indexing DONE_CODES, margin [B]
    = smallest slack of the threshold comparisons evaluated before the row's code was decided)."""
    nxt = np.asarray(nxt, dtype=f32)
    veh = np.asarray(veh_after, dtype=f32).reshape(len(nxt), -1)
    B = len(nxt)
    r, x, y, phi = (nxt[:, i].astype(np.float64) for i in (2, 3, 4, 5))
    v_light = np.broadcast_to(np.asarray(v_light), (B,))
    code = np.zeros(B, np.int8)
    margin = np.full(B, np.inf)

    def decide(cond, slack, c):
        nonlocal code, margin
        open_ = code == 0
        margin = np.where(open_, np.minimum(margin, slack), margin)
        code = np.where(open_ & cond, np.int8(c), code)

    lw = (L - W) / 2                                                                # traffic.py:270-274
    ex0, ey0 = x + np.cos(phi / 180 * np.pi) * lw, y + np.sin(phi / 180 * np.pi) * lw
    ex1, ey1 = x - np.cos(phi / 180 * np.pi) * lw, y - np.sin(phi / 180 * np.pi) * lw
    hit = np.zeros(B, bool)
    slack = np.full(B, np.inf)
    thr = ((W + W) / 2 + 0.5) ** 2                                                  # traffic.py:284
    for j in range(veh.shape[1] // 4):
        vxx, vyy, vph = (veh[:, 4 * j + k].astype(np.float64) for k in (0, 1, 3))
        gate = (np.abs(vxx - x) < 10) & (np.abs(vyy - y) < 10)                      # traffic.py:278
        slack = np.minimum(slack, np.minimum(np.abs(np.abs(vxx - x) - 10), np.abs(np.abs(vyy - y) - 10)))
        sx0, sy0 = vxx + np.cos(vph / 180 * np.pi) * lw, vyy + np.sin(vph / 180 * np.pi) * lw
        sx1, sy1 = vxx - np.cos(vph / 180 * np.pi) * lw, vyy - np.sin(vph / 180 * np.pi) * lw
        for (ax, ay, bx, by) in ((ex0, ey0, sx0, sy0), (ex0, ey0, sx1, sy1), (ex1, ey1, sx1, sy1), (ex1, ey1, sx0, sy0)):
            d2 = (ax - bx) ** 2 + (ay - by) ** 2
            hit |= gate & (d2 < thr)
            slack = np.minimum(slack, np.where(gate, np.abs(d2 - thr), np.inf))
    decide(hit, slack, 1)
    corners, r_bound = gym_ego_dynamics(nxt, params)
    ok = np.ones(B, bool)                                                           # E2E:227-229
    slack = np.full(B, np.inf)
    for k in range(4):
        px, py = corners[:, k, 0], corners[:, k, 1]
        ok &= np.array([judge_feasible(a, b, task) for a, b in zip(px, py)], bool)
        for edge in (-25., 25., 0., LANE_WIDTH, 2 * LANE_WIDTH, 3 * LANE_WIDTH, -3 * LANE_WIDTH):
            slack = np.minimum(slack, np.minimum(np.abs(px - edge), np.abs(py - edge)))
    decide(~ok, slack, 2)
    dy = np.asarray(delta_y).astype(np.float64)
    decide(np.abs(dy) > 15, np.abs(np.abs(dy) - 15), 3)                              # E2E:223-225
    decide(~((-r_bound < r) & (r < r_bound)), np.abs(np.abs(r) - r_bound), 4)       # E2E:231-242
    decide((v_light != 0) & (y > -CROSSROAD_SIZE / 2) & (task != 'right'), np.abs(y + CROSSROAD_SIZE / 2), 5)
    road = LANE_NUMBER * LANE_WIDTH
    if task == 'left':                                                              # E2E:247-256
        goal = (x < -CROSSROAD_SIZE / 2 - 10) & (0 < y) & (y < road)
        slack = np.minimum(np.abs(x + 35), np.minimum(np.abs(y), np.abs(y - road)))
    elif task == 'right':
        goal = (x > CROSSROAD_SIZE / 2 + 10) & (-road < y) & (y < 0)
        slack = np.minimum(np.abs(x - 35), np.minimum(np.abs(y), np.abs(y + road)))
    else:
        goal = (y > CROSSROAD_SIZE / 2 + 10) & (0 < x) & (x < road)
        slack = np.minimum(np.abs(y - 35), np.minimum(np.abs(x), np.abs(x - road)))
    decide(goal, slack, 6)
    return code, margin


def gym_env_step(obs, actions_norm, task, ref_indexes, path_list, mode_list, num_future_data=0, v_light=0):
    """One step of B independent SUMO-free CrossroadEnd2end environments: vectorised restatement of
    endtoend.py:132-144 with the surrounding traffic advanced by the analytic model (veh_predict,
    DM:394-427) in place of Traffic.sim_step.  The ego update, ego dynamics and done logic are
    pinned against the unmodified reference methods (tests/golden/make_golden_env.py ->
    tests/golden/env_*.npz, tests/test_oracle_golden.py); the traffic substitution is this
    project's (the reference needs a SUMO process).

    Returns (next_obs [B,D] f32, reward [B] f32, done_code [B] int8, margin [B] f64)."""
    obs = np.asarray(obs, dtype=f32)
    B = obs.shape[0]
    ntr = 3 * (num_future_data + 1)
    scaled, nxt, params = gym_next_ego_state(obs[:, :6], actions_norm)
    reward = compute_rewards(obs, scaled, task, num_future_data)[0]                 # E2E:501-507
    veh = veh_predict(obs[:, 6 + ntr:], mode_list)                                  # stands in for sim_step
    trk = np.zeros((B, ntr), dtype=f32)
    ref_indexes = np.asarray(ref_indexes)
    for pi in range(len(path_list)):                                                # E2E:293-297
        m = ref_indexes == pi
        if m.any():
            rp = ReferencePath(task, pi, path_list=path_list)
            trk[m] = rp.tracking_error_vector(nxt[m, 3], nxt[m, 4], nxt[m, 5], nxt[m, 0], num_future_data)
    next_obs = np.concatenate([nxt, trk, veh], 1).astype(f32)
    code, margin = gym_judge_done(nxt, params, next_obs[:, 6], veh, task, v_light)
    return next_obs, reward, code, margin


# ----------------------------------------------------------------------------
# interested-vehicle selection -- endtoend.py:340-464
# ----------------------------------------------------------------------------
ROUTE_CLASSES = ('dl', 'du', 'dr', 'rd', 'rl', 'ru', 'ur', 'ud', 'ul', 'lu', 'lr', 'ld')   # E2E:354
VEHICLE_MODE_DICT = dict(left=(('dl', 2), ('du', 2), ('ud', 2), ('ul', 2)),
                         straight=(('dl', 1), ('du', 2), ('ud', 2), ('ru', 2), ('ur', 2)),
                         right=(('dr', 1), ('ur', 2), ('lr', 2)))                           # EU:21-23


def select_interested_vehicles(veh, classes, ego_x, ego_y, task, v_light=0, virtual_red=False):
    """One scene of CrossroadEnd2end._construct_veh_vector_short (E2E:340-464).  veh [N,4] fp32
    rows (x, y, v, phi), classes [N] indexes into ROUTE_CLASSES (other values ignored).
    Returns the [4*VEH_NUM[task]] fp32 vehicle vector.  Pinned against the unmodified reference
    method by tests/golden/make_golden_env.py (sel_* arrays)."""
    half, lw, ln = CROSSROAD_SIZE / 2, LANE_WIDTH, LANE_NUMBER
    ex, ey = f32(ego_x), f32(ego_y)
    lists = {m: [] for m in ROUTE_CLASSES}
    for row, c in zip(np.asarray(veh, dtype=f32), classes):
        if 0 <= c < len(ROUTE_CLASSES):
            lists[ROUTE_CLASSES[c]].append(tuple(float(t) for t in row))
    if task != 'right' and ey < -half and (v_light != 0 or virtual_red):                    # E2E:386-390
        lists['dl'].append((lw / 2, -half + 2.5, 0., 90.))
        lists['du'].append((lw * 1.5, -half + 2.5, 0., 90.))
    X, Y = 0, 1
    keep = dict(                                                                           # E2E:393-411
        dl=lambda v: v[X] > -half - 10 and v[Y] > ey - 2,
        du=lambda v: ey - 2 < v[Y] < half + 10 and v[X] < ex + 5,
        dr=lambda v: v[X] < half + 10 and v[Y] > ey,
        ru=lambda v: v[X] < half + 10 and v[Y] < half + 10,
        ur=(lambda v: v[X] < ex + 7 and ey < v[Y] < half + 10) if task == 'straight' else
           (lambda v: v[X] < half + 10 and v[Y] < half) if task == 'right' else (lambda v: True),
        ud=lambda v: max(ey - 2, -half) < v[Y] < half and ex > v[X],
        ul=lambda v: -half - 10 < v[X] < ex and v[Y] < half,
        lr=lambda v: -half - 10 < v[X] < half + 10)
    order = dict(                                                                          # E2E:414-428
        dl=(lambda v: (v[Y], -v[X]), False), du=(lambda v: v[Y], False), dr=(lambda v: (v[Y], v[X]), False),
        ru=(lambda v: (-v[X], v[Y]), True),
        ur=(lambda v: v[Y], False) if task == 'straight' else (lambda v: (-v[Y], v[X]), True),
        ud=(lambda v: v[Y], False), ul=(lambda v: (-v[Y], -v[X]), True), lr=(lambda v: -v[X], False))
    fill = dict(dl=(lw / 2, -(half + 30), 0, 90), du=(lw * 1.5, -(half + 30), 0, 90),           # E2E:440-447
                dr=(lw * (ln - 0.5), -(half + 30), 0, 90), ru=(half + 15, lw * (ln - 0.5), 0, 180),
                ur=(-lw / 2, half + 20, 0, -90), ud=(-lw * 1.5, half + 20, 0, -90),
                ul=(-lw * (ln - 0.5), half + 20, 0, -90), lr=(-(half + 20), -lw * 1.5, 0, 0))
    out = []
    for mode, num in VEHICLE_MODE_DICT[task]:
        key, rev = order[mode]
        chosen = sorted([v for v in lists[mode] if keep[mode](v)], key=key, reverse=rev)[:num]
        chosen += [fill[mode]] * (num - len(chosen))
        for v in chosen:
            out.extend(v)
    return np.array(out, dtype=f32)


# ------------------------------------------------------------------------------------------
# reset of the batched environment (test infrastructure for ce2e_env_reset)
# ------------------------------------------------------------------------------------------
RESET_SPAN = dict(left=900 + 500, straight=1200 + 500, right=420 + 500)        # endtoend.py:473-478


def reset_init_state(task, path, u_index, u_v):
    """CrossroadEnd2end._reset_init_state (endtoend.py:472-499) with the two `np.random.random()` draws
    passed in: (v_x, v_y, r, x, y, phi) of the fresh ego as float32 (the observation dtype)."""
    random_index = int(u_index * RESET_SPAN[task]) + 700                        # endtoend.py:473-478
    idx = min(max(random_index, 0), len(path[0]) - 1)                           # indexs2points clamp, DM:727-728
    x, y, phi = path[0][idx], path[1][idx], path[2][idx]                        # endtoend.py:480
    v = EXPECTED_V * u_v                                                        # endtoend.py:482
    return np.array([v, 0., 0., x, y, phi], dtype=np.float32)                   # endtoend.py:489-495


def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon et al., SC'11) on arrays: ctr [..., 4], key [..., 2] uint32 -> [..., 4] uint32.
    The counter-based generator behind ce2e_env_reset (csrc/ce2e_rng.h)."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k = [np.asarray(key[..., i], dtype=np.uint64) for i in range(2)]
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), \
        np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k[0], p1 & MASK, (p0 >> np.uint64(32)) ^ c[3] ^ k[1], p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return np.stack(c, axis=-1).astype(np.uint32)


def _u01_24(r):
    return (r >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def env_reset_rows(seed, rows, episodes, task, path_list, V, num_future_data=0, fixed_path=-1):
    """NumPy restatement of k_env_reset (csrc/ce2e.cu): fresh observation rows, path indexes and
    virtual-red-light flags of environments `rows` in their `episodes`-th episode under `seed`.
    Ego state = reset_init_state with u_index = r1 / 2^32, u_v = (r2 >> 8) / 2^24; traffic slots from the
    synthetic distribution of SURVEY 8d; every fp32 operation in the kernel's order."""
    f = np.float32
    rows = np.asarray(rows, dtype=np.uint64)
    ep = np.asarray(episodes, dtype=np.uint64)
    n = len(rows)
    key = np.empty((n, 2), np.uint32)
    key[:, 0], key[:, 1] = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)

    def block(b):
        ctr = np.stack([rows & np.uint64(0xFFFFFFFF), rows >> np.uint64(32), ep, np.full(n, b, np.uint64)], -1)
        return philox4x32_10(ctr.astype(np.uint32), key)

    r = block(0)
    n_paths = len(path_list)
    p = np.full(n, fixed_path, np.int64) if fixed_path >= 0 else \
        ((r[:, 0].astype(np.uint64) * np.uint64(n_paths)) >> np.uint64(32)).astype(np.int64)
    D = 6 + 3 * (num_future_data + 1) + 4 * V
    obs = np.zeros((n, D), f)
    for i in range(n):
        path = path_list[p[i]]
        obs[i, :6] = reset_init_state(task, path, float(r[i, 1]) / 2.0 ** 32, float(_u01_24(r[i:i + 1, 2])[0]))
    trk = np.zeros((n, 3 * (num_future_data + 1)), f)
    for k in range(n_paths):
        m = p == k
        if m.any():
            rp = ReferencePath(task, k, path_list=path_list)
            trk[m] = rp.tracking_error_vector(obs[m, 3], obs[m, 4], obs[m, 5], obs[m, 0], num_future_data)
    obs[:, 6:6 + trk.shape[1]] = trk
    x, y = obs[:, 3], obs[:, 4]
    off = 6 + trk.shape[1]
    for j in range(V):
        a, b = block(1 + 2 * j), block(2 + 2 * j)
        near = (a[:, 0] & np.uint32(0xFFFF)) < np.uint32(6554)
        quad = (a[:, 0] >> np.uint32(16)) & np.uint32(3)
        u1, u2, u3 = _u01_24(a[:, 1]), _u01_24(a[:, 2]), _u01_24(a[:, 3])
        vx = np.where(near, x + (u1 * f(16.) - f(8.)), u1 * f(130.) - f(65.)).astype(f)
        vy = np.where(near, y + (u2 * f(16.) - f(8.)), u2 * f(130.) - f(65.)).astype(f)
        dd = _sq(vx - x) + _sq(vy - y)
        vx = np.where(dd < f(36.), vx + f(30.), vx).astype(f)
        s4 = ((_u01_24(b[:, 0]) + _u01_24(b[:, 1])) + _u01_24(b[:, 2])) + _u01_24(b[:, 3])
        base = np.choose(quad, [f(0.), f(90.), f(180.), f(-90.)]).astype(f)
        vphi = base + f(10.) * ((s4 - f(2.)) * f(1.73205077648162842))
        vphi = np.where(vphi > f(180.), vphi - f(360.), vphi).astype(f)
        vphi = np.where(vphi <= f(-180.), vphi + f(360.), vphi).astype(f)
        obs[:, off + 4 * j] = vx
        obs[:, off + 4 * j + 1] = vy
        obs[:, off + 4 * j + 2] = f(8.) * u3
        obs[:, off + 4 * j + 3] = vphi
    virtual_red = _u01_24(r[:, 3]) > f(0.9)                                    # endtoend.py:120-124
    return obs, p.astype(np.int32), virtual_red
