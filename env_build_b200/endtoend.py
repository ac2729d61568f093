"""Batched, SUMO-free CrossroadEnd2end.

Mirror of the reference's Gym environment (endtoend.py:43-507, E2E below) for the parts that are
arithmetic: `step` = action scaling (E2E:258-267) -> compute_reward on the current observation
(E2E:501-507) -> next ego state (E2E:269-283) -> traffic step -> observation (E2E:285-303) ->
done logic (E2E:200-256).  The reference couples the traffic step to an external SUMO process
(traffic.py), which is out of scope; here the surrounding vehicles follow the analytic model the
reference itself uses for prediction (EnvironmentModel.veh_predict, DM:394-427) and keep their
observation slots, so B environments advance in ONE fused launch plus a small done kernel
(ce2e_env_step).  Same method names and return conventions as the reference: with num_envs == 1
`reset()` returns obs [D] and `step(a[2])` returns (obs [D], reward, done, info) as NumPy / Python
scalars; with num_envs > 1 everything is batched and stays on the device.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import synthetic as syn
from .dynamics_and_models import (EnvironmentModel, ReferencePath, VehicleDynamics, _ptr, _stream, _wrap,
                                  padded_rows, to_device, REWARD_DICT_KEYS)
from .endtoend_env_utils import EXPECTED_V, L, VEH_NUM, VEHICLE_MODE_DICT, VEHICLE_MODE_LIST, W, turn_class

DONE_TYPES = ('not_done_yet', 'collision', 'break_road_constrain', 'deviate_too_much', 'break_stability',
              'break_red_light', 'good_done')
# start-index window of _reset_init_state (E2E:473-478)
_RESET_SPAN = dict(left=900 + 500, straight=1200 + 500, right=420 + 500)


# route classes of the interested-vehicle selection, in the order of the reference's local lists (E2E:354)
ROUTE_CLASSES = ('dl', 'du', 'dr', 'rd', 'rl', 'ru', 'ur', 'ud', 'ul', 'lu', 'lr', 'ld')
# arm names seen from each ego exit (E2E:346-349): d(own) r(ight) u(p) l(eft) -> SUMO edge number
_ARMS = dict(D='1234', R='2341', U='3412', L='4123')


def route_class(route, exit_='D'):
    """Index into ROUTE_CLASSES of a SUMO route (start_edge, end_edge) such as ('1o', '4i') as seen by an
    ego entering from `exit_` (E2E:355-382); -1 for anything else (e.g. the ego's own padding)."""
    try:
        start, end = route[0], route[1]
        arms = _ARMS[exit_]
        a, b = 'drul'[arms.index(start[0])], 'drul'[arms.index(end[0])]
        if start[1] != 'o' or end[1] != 'i' or a == b:
            return -1
        return ROUTE_CLASSES.index(a + b)
    except (ValueError, IndexError, TypeError, KeyError):
        return -1


class Box(object):
    """Minimal stand-in for gym.spaces.Box (gym is not a dependency of this package)."""

    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def sample(self, rng=None):
        rng = rng or np.random
        return rng.uniform(self.low, self.high, self.shape).astype(self.dtype)


class CrossroadEnd2end(object):
    def __init__(self,
                 training_task,  # 'left', 'straight', 'right'
                 num_future_data=0,
                 mode='training',
                 num_envs=1,
                 veh_num=None,
                 auto_reset=False,
                 traffic_init=None,
                 reward_info=True,
                 use_graph=False,
                 **kwargs):
        self.dynamics = VehicleDynamics()
        self.training_task = training_task
        self.num_envs = int(num_envs)
        self.ref_path = ReferencePath(self.training_task, **kwargs)
        self.num_future_data = num_future_data
        self.veh_num = VEH_NUM[training_task] if veh_num is None else int(veh_num)
        self.veh_mode_dict = VEHICLE_MODE_DICT[self.training_task]
        self.veh_mode_list = syn.tiled_mode_list(VEHICLE_MODE_LIST[training_task], self.veh_num)
        self.env_model = EnvironmentModel(training_task, num_future_data, veh_mode_list=self.veh_mode_list)
        self.env_model.ref_path = self.ref_path
        self.action_number = 2
        self.exp_v = EXPECTED_V
        self.ego_l, self.ego_w = L, W
        self.action_space = Box(low=-1, high=1, shape=(self.action_number,), dtype=np.float32)
        self.step_length = 100  # ms
        self.step_time = self.step_length / 1000.0
        self.ego_info_dim, self.per_tracking_info_dim, self.per_veh_info_dim = 6, 3, 4
        self.obs_dim = 6 + 3 * (num_future_data + 1) + 4 * self.veh_num
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(self.obs_dim,), dtype=np.float32)
        self.mode = mode
        self.auto_reset = auto_reset
        # traffic_init(rng, ego_xy, task, V) -> [n, V, 4]: host-side initial traffic (resets then go through
        # the host); None: resets run on the device (ce2e_env_reset), no host round trip
        self.traffic_init = traffic_init
        self.reward_info_enabled = bool(reward_info)     # the 16-term reward dict triples the step's writes
        self.use_graph = bool(use_graph)                 # replay step() as a CUDA graph (batched, device reset)
        self._bufs = None                                # static state, allocated by reset()
        self._graphs, self._act_seen = {}, {}
        self._views, self._cur_obs = {}, None
        self.v_light = 0                     # the model traffic has no signal phases: always green
        self.done_type = 'not_done_yet'
        self.reward_info = None
        self.obs = None
        self.action = None
        self.ref_indexes = None
        self._turn = _lib.make_turn_classes([turn_class(m) for m in self.veh_mode_list])
        self.seed()

    # -- gym plumbing ---------------------------------------------------------------------------
    def seed(self, seed=None):
        self.np_random = np.random.default_rng(seed)
        # key of the device-side draws (ce2e_env_reset: Philox4x32-10 at counter (env, episode, block))
        self._seed = int(seed) & (2 ** 64 - 1) if seed is not None else int(self.np_random.integers(0, 2 ** 63))
        return [seed]

    def close(self):
        self._graphs, self._act_seen = {}, {}
        self.ref_path.close()

    def set_traj(self, trajectory):
        """set the real trajectory to reconstruct observation (E2E:793-795): every environment now
        follows `trajectory` (its ref_index), and the tracking columns of the current observation are
        re-projected onto it, which is what the reference's `env.set_traj(path); env._get_obs()` yields
        (hier_decision.py:115-124, multi_ego.py:104)."""
        self.ref_path = trajectory
        self.env_model.ref_path = trajectory
        self._graphs, self._act_seen = {}, {}                      # captured steps hold the old table handle
        if self.obs is not None:
            self.ref_indexes.fill_(int(trajectory.ref_index))
            self._fill_tracking(self.obs, self.ref_indexes)

    def _squeeze(self, t):
        return t.numpy()[0] if self.num_envs == 1 else t

    # -- reset ------------------------------------------------------------------------------------
    def _reset_rows(self, n):
        """Initial observations of n fresh environments (E2E:472-499 for the ego; the traffic comes
        from `traffic_init(rng, ego_xy, task, V)` -> [n, V, 4] or, by default, from the synthetic
        distribution of env_build_b200.synthetic since no simulator is available)."""
        rng, task, paths = self.np_random, self.training_task, self.ref_path.path_list
        ref = rng.integers(0, len(paths), n).astype(np.int32) if self.num_envs > 1 else \
            np.full(n, self.ref_path.ref_index, np.int32)
        idx = (rng.random(n) * _RESET_SPAN[task]).astype(np.int64) + 700
        obs = np.zeros((n, self.obs_dim), np.float32)
        for p in range(len(paths)):
            m = ref == p
            obs[m, 3], obs[m, 4], obs[m, 5] = paths[p][0][idx[m]], paths[p][1][idx[m]], paths[p][2][idx[m]]
        obs[:, 0] = (EXPECTED_V * rng.random(n)).astype(np.float32)
        if self.veh_num:
            if self.traffic_init is not None:
                veh = np.asarray(self.traffic_init(rng, obs[:, 3:5], task, self.veh_num), np.float32)
            else:
                veh = syn.make_obs(rng, n, task, self.veh_num, paths, ref, 0, edge_frac=0.0, near_frac=0.1)[:, 9:]
                veh = veh.reshape(n, self.veh_num, 4)
                # keep the start collision free: push vehicles inside 6 m of the ego away
                d = np.hypot(veh[:, :, 0] - obs[:, 3:4], veh[:, :, 1] - obs[:, 4:5])
                veh[:, :, 0] = np.where(d < 6, veh[:, :, 0] + 30, veh[:, :, 0])
            obs[:, 6 + 3 * (self.num_future_data + 1):] = veh.reshape(n, -1)
        return obs, ref

    def _fill_tracking(self, obs_dev, ref_dev):
        trk = self.ref_path.tracking_error_vector(obs_dev[:, 3], obs_dev[:, 4], obs_dev[:, 5], obs_dev[:, 0],
                                                  self.num_future_data, ref_indexes=ref_dev)
        obs_dev[:, 6:6 + trk.shape[1]] = trk

    def _alloc(self):
        B, dev = self.num_envs, torch.device('cuda', torch.cuda.current_device())
        veh_off = 6 + 3 * (self.num_future_data + 1)
        f32 = dict(dtype=torch.float32, device=dev)
        self._bufs = dict(obs=[padded_rows(B, self.obs_dim, veh_off, dev) for _ in range(2)],
                          out5=torch.zeros((5, B), **f32),
                          d16=torch.zeros((16, B), **f32) if self.reward_info_enabled else None,
                          scaled=torch.zeros((B, 2), **f32), act=torch.zeros((B, 2), **f32),
                          done=torch.zeros((B,), dtype=torch.int8, device=dev),
                          done_flag=torch.zeros((B,), dtype=torch.bool, device=dev),
                          episode=torch.zeros((B,), dtype=torch.int32, device=dev),
                          ref=torch.zeros((B,), dtype=torch.int32, device=dev),
                          red=torch.zeros((B,), dtype=torch.int8, device=dev))
        for b in self._bufs['obs']:
            b.zero_()
        self._cur = 0
        self._views = {}
        self.action_buffer = self._bufs['act']       # a policy may write its [B, 2] actions here and pass it to step()

    def _device_reset(self, obs, done):
        """ce2e_env_reset on the rows of `obs` whose `done` code is non-zero (done None: all rows)."""
        fixed = -1 if self.num_envs > 1 and self._fixed_path is None else int(
            self.ref_path.ref_index if self._fixed_path is None else self._fixed_path)
        b = self._bufs
        _lib.check(_lib.load().ce2e_env_reset(self.ref_path.handle, ctypes.c_uint64(self._seed), _ptr(b['episode']),
                                              _ptr(done), fixed, _ptr(obs), obs.stride(0), _ptr(b['ref']), _ptr(b['red']),
                                              self.veh_num, int(self.num_future_data), self.num_envs, _stream()))

    def reset(self, **kwargs):
        """E2E:99-127.  Batched environments draw their path per row; `reset(ref_index=k)` (the
        reference's ReferencePath kwargs) pins every row to path k."""
        self._graphs, self._act_seen = {}, {}
        if kwargs:
            self.ref_path = ReferencePath(self.training_task, **kwargs)      # E2E:100
            self.env_model.ref_path = self.ref_path
        elif self.num_envs == 1:                                             # E2E:100 -> DM:591: a fresh random path
            self.ref_path.set_path(int(self.np_random.integers(len(self.ref_path.path_list))))
        self._fixed_path = int(self.ref_path.ref_index) if (kwargs.get('ref_index') is not None or self.num_envs == 1) \
            else None
        if self._bufs is None:
            self._alloc()
        b = self._bufs
        buf = b['obs'][self._cur]
        self.ref_indexes = b['ref']
        if self.traffic_init is None:
            self._device_reset(buf, None)
        else:
            obs, ref = self._reset_rows(self.num_envs)
            buf.copy_(to_device(obs))
            self.ref_indexes.copy_(to_device(ref, torch.int32))
            self._fill_tracking(buf, self.ref_indexes)
        self.obs = self._cur_obs = _wrap(buf)
        self.action = None
        self.reward_info = None
        self.done_type = 'not_done_yet'
        self.virtual_red_light_vehicle = bool(b['red'][0].item()) if (self.num_envs == 1 and self.mode == 'training') \
            else False                                                       # E2E:119-126
        return self._squeeze(self.obs)

    # -- step -------------------------------------------------------------------------------------
    def _enqueue_step(self, cur, act=None):
        """One environment step on the static buffers: the fused model step + done kernel
        (ce2e_env_step) and, with auto_reset on the device path, the reset kernel.  No allocation, no
        host synchronisation: capturable.  `act`: device pointer of the [B, 2] actions (default: the
        action buffer)."""
        b = self._bufs
        obs, nxt = b['obs'][cur], b['obs'][1 - cur]
        act = _ptr(b['act']) if act is None else act
        if self.auto_reset and self.num_envs > 1 and self.traffic_init is None:
            fixed = -1 if self._fixed_path is None else int(self._fixed_path)
            _lib.check(_lib.load().ce2e_env_step_reset(
                self.ref_path.handle, _ptr(b['ref']), _ptr(obs), obs.stride(0), act, ctypes.byref(self._turn),
                self.veh_num, int(self.num_future_data), int(self.v_light), _ptr(nxt), nxt.stride(0), _ptr(b['out5']),
                _ptr(b['d16']), _ptr(b['scaled']), _ptr(b['done']), _ptr(b['done_flag']), ctypes.c_uint64(self._seed),
                _ptr(b['episode']), fixed, _ptr(b['red']), self.num_envs, _stream()))
            return
        _lib.check(_lib.load().ce2e_env_step(self.ref_path.handle, _ptr(b['ref']), _ptr(obs), obs.stride(0),
                                             act, ctypes.byref(self._turn), self.veh_num,
                                             int(self.num_future_data), int(self.v_light), _ptr(nxt), nxt.stride(0),
                                             _ptr(b['out5']), _ptr(b['d16']), _ptr(b['scaled']), _ptr(b['done']),
                                             self.num_envs, _stream()))
        if self.num_envs > 1:
            torch.ne(b['done'], 0, out=b['done_flag'])
        if self.auto_reset and self.num_envs > 1 and self.traffic_init is None:
            self._device_reset(nxt, b['done'])

    def _views_for(self, cur):
        """The wrapped views a batched step returns when the observations are in buffer `cur` (cached: a
        step must not cost tens of microseconds of Python)."""
        v = self._views.get(cur)
        if v is None:
            b = self._bufs
            info = dict(done_code=_wrap(b['done']), ref_index=b['ref'])
            if b['d16'] is not None:
                info['reward_info'] = {k: _wrap(b['d16'][i]) for i, k in enumerate(REWARD_DICT_KEYS)}
            v = dict(obs=_wrap(b['obs'][cur]), scaled=_wrap(b['scaled']), info=info,
                     ret=(_wrap(b['obs'][cur]), _wrap(b['out5'][0]), _wrap(b['done_flag']), info))
            v['ret'] = (v['obs'],) + v['ret'][1:]
            self._views[cur] = v
        return v

    def _capture(self, cur, act=None):
        b = self._bufs
        self.ref_path.handle                   # create the device tables outside the capture
        state = (b['obs'][0], b['obs'][1], b['episode'], b['ref'])
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            snap = [t.clone() for t in state]
            self._enqueue_step(cur, act)       # warm-up outside the capture, then undo its effects
            for t, c in zip(state, snap):
                t.copy_(c)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue_step(cur, act)
        if len(self._graphs) >= 8:             # callers that keep changing their action tensor: start over
            self._graphs.clear()
        self._graphs[(cur, None if act is None else act.value)] = g

    def step(self, action):
        """E2E:132-144.  Batched environments return views of static device buffers (observations,
        rewards, done flags, info tensors): they are overwritten by the next step() -- clone what must
        be kept.  With auto_reset the returned observation rows of finished environments are already
        those of their next episode, and `done` still flags them.  `action` may be `env.action_buffer`
        itself (filled in place by the policy), which saves the copy; a contiguous float32 CUDA tensor is
        read in place too (a graph captured for its address once the same tensor has been passed a few
        times), anything else is copied into the action buffer."""
        B = self.num_envs
        if self._bufs is None:
            raise RuntimeError('call reset() before step()')
        b = self._bufs
        graphed = self.use_graph and B > 1 and self.traffic_init is None
        act_ptr = None                                   # None: the action buffer
        if action is not b['act']:
            act = action if isinstance(action, torch.Tensor) else to_device(np.asarray(action, np.float32))
            in_place = (act.is_cuda and act.dtype == torch.float32 and act.is_contiguous() and act.numel() == 2 * B
                        and act.device == b['act'].device and not act.requires_grad and act.data_ptr() % 8 == 0)
            if in_place and graphed:
                # a graph per action address pays off only for a tensor that keeps coming back
                key = act.data_ptr()
                seen = self._act_seen.get(key, 0) + 1
                if len(self._act_seen) > 64:
                    self._act_seen.clear()
                self._act_seen[key] = seen
                in_place = seen >= 3 or (self._cur, key) in self._graphs or (1 - self._cur, key) in self._graphs
            if in_place:
                act_ptr = ctypes.c_void_p(act.data_ptr())
            else:
                b['act'].copy_(act.reshape(B, 2), non_blocking=True)
        cur = self._cur
        # a caller may have assigned env.obs / env.ref_indexes (the reference's attributes): adopt them
        if self.obs is not self._cur_obs:
            if self.obs is not None and self.obs.data_ptr() != b['obs'][cur].data_ptr():
                b['obs'][cur].copy_(to_device(self.obs).reshape(B, self.obs_dim), non_blocking=True)
        if self.ref_indexes is not b['ref']:
            if self.ref_indexes is not None and self.ref_indexes.data_ptr() != b['ref'].data_ptr():
                b['ref'].copy_(to_device(self.ref_indexes, torch.int32).reshape(B), non_blocking=True)
            self.ref_indexes = b['ref']
        if graphed:
            key = (cur, None if act_ptr is None else act_ptr.value)
            if key not in self._graphs:
                self._capture(cur, act_ptr)
            self._graphs[key].replay()
        else:
            self._enqueue_step(cur, act_ptr)
        self._cur = cur = 1 - cur
        if B > 1:
            v = self._views_for(cur)
            self.action, self.obs, self._cur_obs, self.done_code = v['scaled'], v['obs'], v['obs'], v['info']['done_code']
            if self.auto_reset and self.traffic_init is not None:
                self._reset_done_rows(b['done'])
            return v['ret']
        self.action = _wrap(b['scaled'])
        self.obs = self._cur_obs = _wrap(b['obs'][cur])
        self.done_code = _wrap(b['done'])
        reward, d16 = b['out5'][0], b['d16']
        code = int(b['done'].item())
        self.done_type = DONE_TYPES[code]
        if d16 is not None:
            self.reward_info = {k: float(d16[i, 0]) for i, k in enumerate(REWARD_DICT_KEYS)}
            self.reward_info.update({'final_rew': float(reward[0])})
        info = dict(reward_info=self.reward_info, ref_index=int(self.ref_indexes[0]), done_type=self.done_type,
                    ego_dynamics=self._ego_dynamics_dict())
        return self.obs.numpy()[0], float(reward[0]), int(code != 0), info

    def _reset_done_rows(self, done):
        """Host-side auto-reset, only for a user-supplied `traffic_init` (synchronises)."""
        rows = torch.nonzero(done != 0).reshape(-1)
        n = int(rows.numel())
        if n == 0:
            return
        obs, ref = self._reset_rows(n)
        fresh = to_device(obs)
        ref_dev = to_device(ref, torch.int32)
        self._fill_tracking(fresh, ref_dev)
        self.obs[rows] = fresh
        self.ref_indexes[rows] = ref_dev

    def _judge_done(self, obs=None, scaled_action=None):
        """E2E:200-256 on the current (or given) post-step observations -> (done_type, done) for one
        environment, (codes, done) tensors for a batch."""
        obs = self.obs if obs is None else self.env_model._adopt(np.atleast_2d(np.asarray(obs, np.float32)))
        B = obs.shape[0]
        if scaled_action is not None:
            act = to_device(np.atleast_2d(np.asarray(scaled_action, np.float32)))
        elif self.action is not None:
            act = self.action
        else:                                   # right after reset(): miu_r = miu like E2E:111-113 (a_x = 0)
            act = torch.zeros((B, 2), dtype=torch.float32, device=obs.device)
        done = torch.empty((B,), dtype=torch.int8, device=obs.device)
        _lib.check(_lib.load().ce2e_judge_done(_lib.TASK_ID[self.training_task], _ptr(obs), obs.stride(0) if B > 1 else
                                               max(obs.stride(0), obs.shape[1]), _ptr(act.contiguous()), self.veh_num,
                                               int(self.num_future_data), int(self.v_light), _ptr(done), B, _stream()))
        if B == 1:
            code = int(done.item())
            return DONE_TYPES[code], int(code != 0)
        return _wrap(done), _wrap(done != 0)

    # -- interested-vehicle selection (E2E:340-464) ------------------------------------------------
    def construct_veh_vectors(self, veh_all, route_classes, ego_xy, virtual_red=None):
        """Batched _construct_veh_vector_short: veh_all [B,N,4] (x, y, v, phi), route_classes [B,N]
        (indexes into ROUTE_CLASSES, see route_class()), ego_xy [B,2] -> [B, 4*VEH_NUM[task]]."""
        veh = to_device(veh_all).contiguous()
        cls = to_device(route_classes, torch.int8).contiguous()
        ego = to_device(ego_xy).contiguous()
        B, N = veh.shape[0], veh.shape[1]
        if veh.dim() != 3 or veh.shape[2] != 4 or tuple(cls.shape) != (B, N) or tuple(ego.shape) != (B, 2):
            raise ValueError('expected veh_all [B,N,4], route_classes [B,N], ego_xy [B,2]')
        vr = None if virtual_red is None else to_device(virtual_red, torch.int8).reshape(B).contiguous()
        V = VEH_NUM[self.training_task]
        out = torch.empty((B, 4 * V), dtype=torch.float32, device=veh.device)
        _lib.check(_lib.load().ce2e_select_vehicles(_lib.TASK_ID[self.training_task], _ptr(veh), _ptr(cls), N, _ptr(ego),
                                                    int(self.v_light), _ptr(vr), _ptr(out), 4 * V, B, _stream()))
        return _wrap(out)

    def _construct_veh_vector_short(self, exit_='D'):
        """E2E:340-464 for one environment whose `all_vehicles` is the reference's list of dicts
        (x, y, v, phi, route, ...) and whose ego position is taken from the current observation."""
        vehs = getattr(self, 'all_vehicles', None) or []
        veh = np.array([[v['x'], v['y'], v['v'], v['phi']] for v in vehs], np.float32).reshape(1, -1, 4)
        cls = np.array([route_class(v.get('route'), exit_) for v in vehs], np.int8).reshape(1, -1)
        ego = self.obs.numpy()[:1, 3:5]
        vr = np.array([1 if getattr(self, 'virtual_red_light_vehicle', False) else 0], np.int8)
        return self.construct_veh_vectors(veh, cls, ego, vr).numpy()[0]

    # -- reference-named numeric helpers (batch-1 callers) ---------------------------------------
    def _ego_dynamics_dict(self):
        o = self.obs.numpy()[0]
        return dict(v_x=o[0], v_y=o[1], r=o[2], x=o[3], y=o[4], phi=o[5], l=self.ego_l, w=self.ego_w)

    def _action_transformation_for_end2end(self, action):
        a = self.env_model._action_transformation_for_end2end(np.asarray(action, np.float32).reshape(-1, 2))
        return a.numpy()[0] if np.ndim(action) == 1 else a

    def compute_reward(self, obs, action):
        """E2E:501-507: `action` is the scaled action; returns (reward, reward_dict) of one row."""
        res = self.env_model.compute_rewards(np.asarray(obs, np.float32)[np.newaxis, :],
                                             np.asarray(action, np.float32)[np.newaxis, :])
        return res[0].numpy()[0], {k: v.numpy()[0] for k, v in res[5].items()}

    def _get_next_ego_state(self, trans_action):
        """E2E:269-283 for the current (single) environment."""
        state = self.obs[:, :6] if self.obs.dim() == 2 else self.obs[None, :6]
        nxt, par = self.dynamics.prediction(state, np.asarray(trans_action, np.float32).reshape(-1, 2), 10)
        nxt, par = nxt.numpy(), par.numpy()
        nxt[:, 0] = np.where(nxt[:, 0] >= 0, nxt[:, 0], 0.)
        from .endtoend_env_utils import deal_with_phi
        nxt[:, 5] = [deal_with_phi(float(p)) for p in nxt[:, 5]]
        return (nxt[0], par[0]) if self.num_envs == 1 else (nxt, par)

    def _get_obs(self, exit_='D'):
        """E2E:285-303: the observation of the current state; its tracking columns are projected onto the
        CURRENT reference path (self.ref_path / the per-row path indexes), like the reference's."""
        self._fill_tracking(self.obs, self.ref_indexes)
        return self._squeeze(self.obs)

    def render(self, mode='human'):
        raise NotImplementedError('rendering is out of scope (SURVEY.md section 2)')
