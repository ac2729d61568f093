"""Host-side mirror of the reference's dynamics_and_models.py for the model hot path.

Same class names, method names, argument meaning and error behaviour as the reference
(VehicleDynamics DM:26-87, EnvironmentModel DM:90-427, ReferencePath DM:583-770,
deal_with_phi_diff DM:577-580; DM = reference dynamics_and_models.py), but every tensor
operation runs in libce2e.so's sm_100a kernels through the C ABI of include/ce2e.h.
Inputs may be NumPy arrays, Python lists or torch tensors (host or device); outputs are
CUDA tensors (`DeviceTensor`, whose .numpy() copies to the host like an EagerTensor's).
There is no CPU path: without a CUDA device or without the built library every call raises.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib
from .endtoend_env_utils import (CROSSROAD_SIZE, EXPECTED_V, LANE_NUMBER, LANE_WIDTH, TASKS,
                                 VEHICLE_MODE_LIST, turn_class)

REWARD_DICT_KEYS = ('punish_steer', 'punish_a_x', 'punish_yaw_rate', 'devi_v', 'devi_y', 'devi_phi',
                    'scaled_punish_steer', 'scaled_punish_a_x', 'scaled_punish_yaw_rate', 'scaled_devi_v',
                    'scaled_devi_y', 'scaled_devi_phi', 'veh2veh4training', 'veh2road4training',
                    'veh2veh4real', 'veh2road4real')


# ------------------------------------------------------------------------------------------
# tensors
# ------------------------------------------------------------------------------------------
class DeviceTensor(torch.Tensor):
    """A CUDA tensor whose .numpy() works (callers of the reference do `.numpy()[0]` on
    results, e.g. endtoend.py:280, :297, :506)."""

    def numpy(self, *args, **kwargs):
        return self.detach().as_subclass(torch.Tensor).cpu().numpy(*args, **kwargs)


def _wrap(t):
    return t.as_subclass(DeviceTensor)


_CUDA_CHECKED = False
_RAW_STREAM = getattr(torch._C, '_cuda_getCurrentRawStream', None)

# NVTX ranges named after the reference's tf.name_scope labels ('model_step' DM:119, 'compute_reward'
# DM:188) around the corresponding launches, for nsys / ncu --nvtx timelines.  Off unless CE2E_NVTX=1.
_NVTX = os.environ.get('CE2E_NVTX') == '1'


class _nvtx_range(object):
    __slots__ = ('name',)

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def _device():
    global _CUDA_CHECKED
    if not _CUDA_CHECKED:
        if not torch.cuda.is_available():
            raise RuntimeError('env_build_b200 needs a CUDA device: the model path has no CPU fallback')
        _CUDA_CHECKED = True
    return torch.device('cuda', torch.cuda.current_device())


def _raw(t):
    """A plain torch.Tensor view of `t` (DeviceTensor's __torch_function__ costs microseconds per op)."""
    return t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t


def to_device(x, dtype=torch.float32):
    """NumPy / list / torch (any device) -> CUDA tensor of `dtype` (no copy if already there)."""
    if isinstance(x, torch.Tensor):
        t = _raw(x)
        # the launch stream and the path-table handle belong to the CURRENT device: a tensor living on
        # another GPU is moved there (a silent peer access or an illegal address otherwise)
        if t.is_cuda and t.dtype == dtype and t.device.index == torch.cuda.current_device():
            return t
        return t.to(device=_device(), dtype=dtype)
    return torch.as_tensor(np.asarray(x), device=_device()).to(dtype)


def _stream():
    if _RAW_STREAM is not None:
        return ctypes.c_void_p(_RAW_STREAM(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _rows(t, what, cols=None):
    if t.dim() != 2 or (cols is not None and t.shape[1] != cols):
        raise ValueError('%s must have shape [B, %s], got %s' % (what, cols if cols else 'D', tuple(t.shape)))
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def _ld(t):
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


def padded_rows(B, D, veh_off, device=None):
    """An uninitialised [B, D] fp32 view in the row layout the kernels like: the first vehicle column
    (`veh_off`) of every row is 16-byte aligned (float4 loads, TMA boxes; 64 B for row 0), the row stride is a
    multiple of 4 floats, and with veh_off == 9 (no preview points) every row is preceded by at least 7
    floats of padding, so that the 64 B in front of a row's vehicle block (its ego + tracking columns) can
    travel as one TMA box without touching the previous row's data.  V = 32: 144 floats per row, V = 8 / 9 /
    5: 48 / 52 / 36."""
    front = (-veh_off) % 16
    ld = -(-(D + (7 if veh_off == 9 else 0)) // 4) * 4
    store = torch.empty(max(B, 1) * ld + 16, dtype=torch.float32, device=device or _device())
    return store.as_strided((B, D), (ld, 1), front)


def deal_with_phi_diff(phi_diff):
    """DM:577-580 (one wrap each side) on tensors / arrays."""
    t = to_device(phi_diff)
    t = torch.where(t > 180., t - 360., t)
    t = torch.where(t < -180., t + 360., t)
    return _wrap(t)


# ------------------------------------------------------------------------------------------
# VehicleDynamics
# ------------------------------------------------------------------------------------------
class VehicleDynamics(object):
    """DM:26-87."""

    def __init__(self, ):
        # single-track model parameters, values of DM:37-44 (SI units): tyre cornering stiffness front /
        # rear (N/rad), axle distances from the centre of gravity (m), mass (kg), yaw inertia (kg m^2),
        # tyre-road friction, gravity (m/s^2)
        self.vehicle_params = dict(C_f=-155495.0, C_r=-155495.0, a=1.19, b=1.46, mass=1520., I_z=2642.,
                                   miu=0.8, g=9.81)
        a, b, mass, g = (self.vehicle_params[k] for k in ('a', 'b', 'mass', 'g'))
        self.vehicle_params.update(dict(F_zf=b * mass * g / (a + b), F_zr=a * mass * g / (a + b)))

    def f_xu(self, states, actions, tau, clip_vx=False):
        """states [B,6] (phi in degrees), actions [B,2] = (steer rad, a_x), tau seconds ->
        (next_states [B,6], [alpha_f, alpha_r, miu_f, miu_r] [B,4])."""
        st = _rows(to_device(states), 'states', 6)
        ac = _rows(to_device(actions), 'actions', 2).contiguous()
        B = st.shape[0]
        if ac.shape[0] != B:
            raise ValueError('states and actions disagree on the batch size')
        nxt = torch.empty((B, 6), dtype=torch.float32, device=st.device)
        par = torch.empty((B, 4), dtype=torch.float32, device=st.device)
        _lib.check(_lib.load().ce2e_dynamics_step(_ptr(st), _ld(st), _ptr(ac), float(tau), _ptr(nxt), 6,
                                                  _ptr(par), int(bool(clip_vx)), B, _stream()))
        return _wrap(nxt), _wrap(par)

    def prediction(self, x_1, u_1, frequency):
        return self.f_xu(x_1, u_1, 1 / frequency)


# ------------------------------------------------------------------------------------------
# ReferencePath
# ------------------------------------------------------------------------------------------
def _cubic_bezier(ctrl_xy32, n):
    """n points of the cubic with float32-rounded control points (DM:613-619), evaluated in
    float64 the way bezier.Curve.evaluate_multi does (barycentric Horner form)."""
    P = np.asarray(ctrl_xy32, dtype=np.float32).astype(np.float64)       # [4, 2]
    t = np.linspace(0., 1., n)
    u = 1. - t
    out = []
    for d in (0, 1):
        acc = u * P[0, d]
        acc = (acc + (3. * t) * P[1, d]) * u
        tt = t * t
        acc = (acc + (3. * tt) * P[2, d]) * u
        acc = acc + (t * tt) * P[3, d]
        out.append(acc.astype(np.float32))
    return out


def build_path_tables(task):
    """The task's three static reference paths (DM:598-700): 40 m approach line, cubic Bezier
    through the junction, 40 m exit line, 30 points per metre; heading of each point is the
    direction to its successor in degrees.  Returns (path_list, path_len_list, control_points)."""
    if task not in TASKS:
        raise AssertionError('task must be one of %s' % (TASKS,))
    from .endtoend_env_utils import get_config
    cfg = get_config()                       # the reference reads module constants (EU:14-18)
    CROSSROAD_SIZE, LANE_WIDTH, LANE_NUMBER = cfg.CROSSROAD_SIZE, cfg.LANE_WIDTH, cfg.LANE_NUMBER
    half = CROSSROAD_SIZE / 2
    line_m, per_m = 40, 30
    n_line = line_m * per_m
    lane_mid = [LANE_WIDTH * (i + 0.5) for i in range(LANE_NUMBER)]
    if task == 'left':
        start_x, ext = lane_mid[0], CROSSROAD_SIZE / 3.
        exits = [dict(ctrl=((-half + ext, off), (-half, off)), axis='x', sign=-1, fixed=off) for off in lane_mid]
        n_curve = int(math.pi / 2 * (half + LANE_WIDTH / 2)) * per_m
    elif task == 'straight':
        start_x, ext = lane_mid[1], CROSSROAD_SIZE / 3.
        exits = [dict(ctrl=((off, half - ext), (off, half)), axis='y', sign=+1, fixed=off) for off in lane_mid]
        n_curve = CROSSROAD_SIZE * per_m
    else:
        start_x, ext = lane_mid[2], CROSSROAD_SIZE / 5.
        exits = [dict(ctrl=((half - ext, -off), (half, -off)), axis='x', sign=+1, fixed=-off)
                 for off in reversed(lane_mid)]
        n_curve = int(math.pi / 2 * (half - lane_mid[2])) * per_m
    f32 = np.float32
    paths, lens, ctrls = [], [], []
    for ex in exits:
        ctrl = [(start_x, -half), (start_x, -half + ext), ex['ctrl'][0], ex['ctrl'][1]]
        ctrls.append(ctrl)
        cx, cy = _cubic_bezier(ctrl, n_curve)
        in_x = np.full(n_line - 1, start_x, dtype=f32)
        in_y = np.linspace(-half - line_m, -half, n_line, dtype=f32)[:-1]
        ramp = np.linspace(half, half + line_m, n_line, dtype=f32)[1:]
        flat = np.full(n_line - 1, ex['fixed'], dtype=f32)
        if ex['axis'] == 'x':
            out_x, out_y = (ramp if ex['sign'] > 0 else np.linspace(-half, -half - line_m, n_line, dtype=f32)[1:]), flat
        else:
            out_x, out_y = flat, ramp
        xs = np.concatenate([in_x, cx, out_x])
        ys = np.concatenate([in_y, cy, out_y])
        ang = np.arctan2((ys[1:] - ys[:-1]).astype(np.float64), (xs[1:] - xs[:-1]).astype(np.float64)).astype(f32)
        phis = ang * f32(180) / f32(math.pi)
        paths.append((xs[:-1].copy(), ys[:-1].copy(), phis))
        lens.append((n_line, n_curve, len(phis)))
    return paths, lens, ctrls


class ReferencePath(object):
    """DM:583-770.  `path_list` holds host float32 arrays like the reference; the device copy
    is created on first use and owned by this object."""

    def __init__(self, task, ref_index=None, path_list=None):
        self.exp_v = EXPECTED_V
        self.task = task
        if path_list is None:
            self.path_list, self.path_len_list, self.control_points = build_path_tables(task)
        else:
            if task not in TASKS:
                raise AssertionError('task must be one of %s' % (TASKS,))
            self.path_list = [tuple(np.ascontiguousarray(a, dtype=np.float32) for a in p) for p in path_list]
            self.path_len_list, self.control_points = [], []
        self.ref_index = np.random.choice(len(self.path_list)) if ref_index is None else ref_index
        self._handles = {}                    # CUDA device ordinal -> ce2e_paths handle

    # -- path selection -------------------------------------------------------------------
    @property
    def path(self):
        return self.path_list[self.ref_index]

    @path.setter
    def path(self, value):
        for i, p in enumerate(self.path_list):
            if p is value:
                self.ref_index = i
                return
        raise ValueError('path must be an element of path_list (use set_path)')

    def set_path(self, path_index=None):
        self.path_list[path_index]            # IndexError / TypeError like the reference
        self.ref_index = path_index

    # -- device tables --------------------------------------------------------------------
    @property
    def handle(self):
        """The device-resident tables on the CURRENT CUDA device (created on first use, one per device)."""
        dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
        h = self._handles.get(dev)
        if h is None:
            _device()
            n = len(self.path_list)
            if n > _lib.MAX_PATHS:
                raise ValueError('at most %d paths' % _lib.MAX_PATHS)
            lens = (ctypes.c_int32 * n)(*[len(p[0]) for p in self.path_list])
            cols = []
            for c in range(3):
                cols.append((ctypes.c_void_p * n)(*[p[c].ctypes.data for p in self.path_list]))
            h = ctypes.c_void_p()
            _lib.check(_lib.load().ce2e_paths_create(_lib.TASK_ID[self.task], n, lens, cols[0], cols[1], cols[2],
                                                     ctypes.byref(h)))
            self._handles[dev] = h
        return h

    def close(self):
        handles, self._handles = getattr(self, '_handles', {}), {}
        for h in handles.values():
            _lib.load().ce2e_paths_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- queries --------------------------------------------------------------------------
    def find_closest_point(self, xs, ys, ratio=10, brute_force=False):
        xs, ys = to_device(xs).reshape(-1).contiguous(), to_device(ys).reshape(-1).contiguous()
        B = xs.shape[0]
        idx = torch.empty((B,), dtype=torch.int64, device=xs.device)
        pts = torch.empty((3, B), dtype=torch.float32, device=xs.device)
        _lib.check(_lib.load().ce2e_find_closest_point(self.handle, int(self.ref_index), _ptr(xs), _ptr(ys),
                                                       int(ratio), int(bool(brute_force)), _ptr(idx), _ptr(pts), B,
                                                       _stream()))
        return _wrap(idx), (_wrap(pts[0]), _wrap(pts[1]), _wrap(pts[2]))

    def future_n_data(self, current_indexs, n):
        idx = to_device(current_indexs, torch.int64).reshape(-1).contiguous()
        B = idx.shape[0]
        if n == 0:
            return []
        pts = torch.empty((n, 3, B), dtype=torch.float32, device=idx.device)
        _lib.check(_lib.load().ce2e_index_points(self.handle, int(self.ref_index), _ptr(idx), int(n), _ptr(pts),
                                                 B, _stream()))
        return [(_wrap(pts[k, 0]), _wrap(pts[k, 1]), _wrap(pts[k, 2])) for k in range(n)]

    def indexs2points(self, indexs):
        idx = to_device(indexs, torch.int64).reshape(-1).contiguous()
        B = idx.shape[0]
        pts = torch.empty((3, B), dtype=torch.float32, device=idx.device)
        _lib.check(_lib.load().ce2e_index_points(self.handle, int(self.ref_index), _ptr(idx), 0, _ptr(pts), B,
                                                 _stream()))
        return _wrap(pts[0]), _wrap(pts[1]), _wrap(pts[2])

    def tracking_error_vector(self, ego_xs, ego_ys, ego_phis, ego_vs, n, ref_indexes=None):
        """DM:735-770 -> [B, 3(n+1)].  `ref_indexes` ([B] ints, extension) selects a path per
        row like EnvironmentModel's training mode; default is the current path."""
        cols = [to_device(a).reshape(-1).contiguous() for a in (ego_xs, ego_ys, ego_phis, ego_vs)]
        B = cols[0].shape[0]
        if any(c.shape[0] != B for c in cols):
            raise ValueError('ego_xs, ego_ys, ego_phis, ego_vs must have the same length')
        ref = None if ref_indexes is None else to_device(ref_indexes, torch.int32).reshape(-1).contiguous()
        out = torch.empty((B, 3 * (n + 1)), dtype=torch.float32, device=cols[0].device)
        _lib.check(_lib.load().ce2e_tracking_error(self.handle, int(self.ref_index), _ptr(ref), _ptr(cols[0]),
                                                   _ptr(cols[1]), _ptr(cols[2]), _ptr(cols[3]), int(n),
                                                   _ptr(out), 3 * (n + 1), B, _stream()))
        return _wrap(out)


# ------------------------------------------------------------------------------------------
# EnvironmentModel
# ------------------------------------------------------------------------------------------
class EnvironmentModel(object):  # all tensors
    """DM:90-427 (render excluded).  `veh_mode_list` (extension) overrides
    VEHICLE_MODE_LIST[task]; it is what makes observations with V != VEH_NUM[task]
    vehicles (e.g. the V=32 benchmark rows) well defined."""

    def __init__(self, training_task, num_future_data=0, mode='training', veh_mode_list=None):
        self.task = training_task
        self.mode = mode
        self.vehicle_dynamics = VehicleDynamics()
        self.base_frequency = 10.
        self._obs = None                 # raw padded device tensor behind the `obses` attribute
        self._ref = None
        self.ego_params = None
        self.actions = None
        self.ref_path = ReferencePath(self.task)
        self.num_future_data = num_future_data
        self.exp_v = EXPECTED_V
        self.reward_info = None
        self.ego_info_dim = 6
        self.per_veh_info_dim = 4
        self.per_tracking_info_dim = 3
        self.set_veh_mode_list(VEHICLE_MODE_LIST[self.task] if veh_mode_list is None else veh_mode_list)

    # -- plumbing ---------------------------------------------------------------------------
    def set_veh_mode_list(self, modes):
        self.veh_mode_list = list(modes)
        self._turn = _lib.make_turn_classes([turn_class(m) for m in self.veh_mode_list])
        self._turn_ref = ctypes.byref(self._turn)

    @property
    def obses(self):
        """The current observations (DM:100), a CUDA tensor in the padded row layout."""
        return None if self._obs is None else _wrap(self._obs)

    @obses.setter
    def obses(self, value):
        self._obs = None if value is None else self._adopt(value)

    @property
    def ref_indexes(self):
        return None if self._ref is None else _wrap(self._ref)

    @ref_indexes.setter
    def ref_indexes(self, value):
        self._ref = None if value is None else to_device(value, torch.int32).reshape(-1).contiguous()

    @property
    def _veh_off(self):
        return self.ego_info_dim + self.per_tracking_info_dim * (self.num_future_data + 1)

    def _adopt(self, obses):
        """Bring observations to the device in the padded row layout the kernels like."""
        t = to_device(obses)
        if t.dim() != 2:
            raise ValueError('obses must have shape [B, D], got %s' % (tuple(t.shape),))
        B, D = t.shape
        if D < self._veh_off or (D - self._veh_off) % self.per_veh_info_dim:
            raise ValueError('obses have %d columns; expected %d + 4*V' % (D, self._veh_off))
        aligned = (t.stride(1) == 1 and (B <= 1 or t.stride(0) % 4 == 0) and
                   (t.data_ptr() + 4 * self._veh_off) % 16 == 0)
        if aligned:
            return t
        buf = padded_rows(B, D, self._veh_off, t.device)
        buf.copy_(t)
        return buf

    def _num_veh(self, obses):
        return (obses.shape[1] - self._veh_off) // self.per_veh_info_dim

    def reset(self, obses, ref_indexes=None):
        self.obses = obses
        self.ref_indexes = ref_indexes
        self.actions = None
        self.reward_info = None

    def add_traj(self, obses, path_index):
        self.obses = obses
        self.ref_path.set_path(path_index)

    def _path_args(self, B):
        if self.mode != 'training':
            return int(self.ref_path.ref_index), None
        ref = self._ref
        if ref is None:
            raise ValueError("mode='training' needs per-row ref_indexes (reset(obses, ref_indexes))")
        if ref.shape[0] != B:
            raise ValueError('ref_indexes has %d entries for %d rows' % (ref.shape[0], B))
        return 0, ref

    # -- the hot call -----------------------------------------------------------------------
    def rollout_out(self, actions):
        """DM:118-126: one fused launch (ce2e_rollout_step)."""
        obs = self._obs
        B, D = obs.shape
        act = to_device(actions)
        if act.dim() != 2 or act.shape[1] != 2 or act.shape[0] != B:
            raise ValueError('actions must have shape [%d, 2], got %s' % (B, tuple(act.shape)))
        if not act.is_contiguous():
            act = act.contiguous()
        veh_off = self._veh_off
        V_in, V_out = (D - veh_off) // 4, len(self.veh_mode_list)
        if V_out > V_in:
            raise ValueError('observations hold %d vehicles but the mode list has %d' % (V_in, V_out))
        path_index, ref = self._path_args(B)
        if act.requires_grad or obs.requires_grad:              # differentiable step (autograd.py)
            from .autograd import RolloutStep
            nxt, out5, scaled = RolloutStep.apply(obs, act, self, path_index, ref, V_in, V_out)
            self.actions, self._obs, self.last_out5 = scaled, nxt, out5
            return (nxt,) + tuple(out5.unbind(0))
        dev = obs.device
        nxt = padded_rows(B, veh_off + 4 * V_out, veh_off, dev)
        out5 = torch.empty((5, B), dtype=torch.float32, device=dev)
        scaled = torch.empty((B, 2), dtype=torch.float32, device=dev)
        with _nvtx_range('model_step'):
            rc = _lib.load().ce2e_rollout_step(self.ref_path.handle, path_index, _ptr(ref), obs.data_ptr(), _ld(obs),
                                               act.data_ptr(), self._turn_ref, V_in, V_out, int(self.num_future_data),
                                               nxt.data_ptr(), _ld(nxt), out5.data_ptr(), scaled.data_ptr(), B, _stream())
        if rc:
            _lib.check(rc)
        self.actions = _wrap(scaled)
        self._obs = nxt
        self.last_out5 = out5            # [5, B]: the five returned vectors as one tensor
        return (_wrap(nxt),) + tuple(_wrap(t) for t in out5.unbind(0))

    def candidate_observations(self, obses):
        """The multi-path evaluation pattern of the reference's online decision step
        (hier_decision.py:112-119, multi_ego.py:98-114): every row is repeated once per reference
        path with its tracking columns re-projected onto that path (what `env.set_traj(path);
        env._get_obs()` yields, E2E:285-303).  Returns (obs [P*B, D] ordered path-major, ref_indexes
        [P*B]); feed them to `reset(obs, ref_indexes)` with mode='training' to roll all candidates
        out in one launch per step."""
        obs = self._adopt(obses)
        B, D = obs.shape
        P = len(self.ref_path.path_list)
        out = padded_rows(P * B, D, self._veh_off, obs.device)
        ref = torch.arange(P, dtype=torch.int32, device=obs.device).repeat_interleave(B)
        for p in range(P):
            out[p * B:(p + 1) * B] = obs
        trk = self.ref_path.tracking_error_vector(out[:, 3], out[:, 4], out[:, 5], out[:, 0], self.num_future_data,
                                                  ref_indexes=ref)
        out[:, 6:6 + trk.shape[1]] = _raw(trk)
        return _wrap(out), _wrap(ref)

    def _action_transformation_for_end2end(self, actions):
        act = _rows(to_device(actions), 'actions', 2).contiguous()
        out = torch.empty_like(act)
        _lib.check(_lib.load().ce2e_action_transform(_ptr(act), _ptr(out), act.shape[0], _stream()))
        return _wrap(out)

    def compute_rewards(self, obses, actions):
        """DM:186-320; `actions` are the scaled actions."""
        obs = self._adopt(obses)
        B = obs.shape[0]
        act = _rows(to_device(actions), 'actions', 2).contiguous()
        if act.shape[0] != B:
            raise ValueError('actions have %d rows for %d observations' % (act.shape[0], B))
        out5 = torch.empty((5, B), dtype=torch.float32, device=obs.device)
        d16 = torch.empty((16, B), dtype=torch.float32, device=obs.device)
        with _nvtx_range('compute_reward'):
            _lib.check(_lib.load().ce2e_compute_rewards(_lib.TASK_ID[self.task], _ptr(obs), _ld(obs), _ptr(act),
                                                        self._num_veh(obs), int(self.num_future_data), _ptr(out5),
                                                        _ptr(d16), B, _stream()))
        reward_dict = {k: _wrap(d16[i]) for i, k in enumerate(REWARD_DICT_KEYS)}
        return tuple(_wrap(out5[i]) for i in range(5)) + (reward_dict,)

    def compute_next_obses(self, obses, actions):
        """DM:322-358; `actions` are the scaled actions."""
        obs = self._adopt(obses)
        B = obs.shape[0]
        act = _rows(to_device(actions), 'actions', 2).contiguous()
        V_in, V_out = self._num_veh(obs), len(self.veh_mode_list)
        if V_out > V_in:
            raise ValueError('observations hold %d vehicles but the mode list has %d' % (V_in, V_out))
        path_index, ref = self._path_args(B)
        nxt = padded_rows(B, self._veh_off + 4 * V_out, self._veh_off, obs.device)
        _lib.check(_lib.load().ce2e_compute_next_obses(self.ref_path.handle, path_index, _ptr(ref), _ptr(obs),
                                                       _ld(obs), _ptr(act), ctypes.byref(self._turn), V_in, V_out,
                                                       int(self.num_future_data), _ptr(nxt), _ld(nxt), B,
                                                       _stream()))
        return _wrap(nxt)

    def ss(self, obses, actions, lam=0.1):
        """DM:134-184: discrete barrier-function penalty; `actions` are normalised."""
        obs = self._adopt(obses)
        scaled = self._action_transformation_for_end2end(actions)
        nxt = self.compute_next_obses(obs, scaled)
        V = self._num_veh(nxt)
        B = obs.shape[0]
        out = torch.empty((B,), dtype=torch.float32, device=obs.device)
        _lib.check(_lib.load().ce2e_ss(_ptr(obs), _ld(obs), _ptr(nxt), _ld(nxt), V, int(self.num_future_data),
                                       float(lam), _ptr(out), B, _stream()))
        return _wrap(out)

    def ego_predict(self, ego_infos, actions):
        """DM:386-392: f_xu at 10 Hz, then v_x clipped to [0, 35]."""
        ego = to_device(ego_infos)[:, :6]
        nxt, _ = self.vehicle_dynamics.f_xu(ego, actions, 1 / self.base_frequency, clip_vx=True)
        return nxt

    def veh_predict(self, veh_infos):
        """DM:394-403: only the first len(veh_mode_list) vehicles are predicted (and returned)."""
        veh = _rows(to_device(veh_infos), 'veh_infos')
        V = len(self.veh_mode_list)
        if veh.shape[1] < 4 * V:
            raise ValueError('veh_infos has %d columns, the mode list needs %d' % (veh.shape[1], 4 * V))
        B = veh.shape[0]
        out = torch.empty((B, 4 * V), dtype=torch.float32, device=veh.device)
        _lib.check(_lib.load().ce2e_veh_predict(_ptr(veh), _ld(veh), ctypes.byref(self._turn), V, _ptr(out),
                                                4 * V, B, _stream()))
        return _wrap(out)

    def predict_for_a_mode(self, vehs, mode):
        """DM:405-427 for one vehicle slot [B,4]."""
        veh = _rows(to_device(vehs), 'vehs', 4)
        B = veh.shape[0]
        out = torch.empty((B, 4), dtype=torch.float32, device=veh.device)
        turn = _lib.make_turn_classes([turn_class(mode)])
        _lib.check(_lib.load().ce2e_veh_predict(_ptr(veh), _ld(veh), ctypes.byref(turn), 1, _ptr(out), 4, B,
                                                _stream()))
        return _wrap(out)
