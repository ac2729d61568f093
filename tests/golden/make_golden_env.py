#!/usr/bin/env python3
"""Golden vectors for the Gym-side numerics, produced by the UNMODIFIED reference methods.

    python tests/golden/make_golden_env.py [--reference /root/reference]   (build container only)

`endtoend.py` and `traffic.py` are imported from the reference tree as they are, on import-only
stand-ins for gym / traci / sumolib (tests/golden/_refshim_env) and the NumPy-backed TensorFlow /
bezier stand-ins of make_golden.py.  `CrossroadEnd2end.__init__` and `Traffic.__init__` start SUMO,
so the objects are created with `object.__new__` and given exactly the attributes the called
methods read; the methods themselves run unmodified:

  CrossroadEnd2end._action_transformation_for_end2end  E2E:258-267
  CrossroadEnd2end._get_next_ego_state                  E2E:269-283
  CrossroadEnd2end._get_ego_dynamics                    E2E:150-183   (corner points, r_bound)
  CrossroadEnd2end.compute_reward                       E2E:501-507
  CrossroadEnd2end._judge_done (+ the five predicates)  E2E:200-256
  Traffic.collision_check                               traffic.py:263-295
  CrossroadEnd2end._construct_veh_vector_short          E2E:340-464
  CrossroadEnd2end._reset_init_state                    E2E:472-499

Output: tests/golden/env_<task>.npz (inputs and outputs of those calls, one row per sample).
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DONE = ('not_done_yet', 'collision', 'break_road_constrain', 'deviate_too_much', 'break_stability',
        'break_red_light', 'good_done')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    args = ap.parse_args()
    os.environ.setdefault('SUMO_HOME', '/nonexistent-sumo')
    sys.path.insert(0, os.path.join(HERE, '_refshim_env'))
    sys.path.insert(0, os.path.join(HERE, '_refshim'))
    sys.path.insert(0, args.reference)
    sys.path.insert(0, ROOT)
    import endtoend as e2e            # the reference's module
    import traffic as tr              # the reference's module
    import dynamics_and_models as dm
    import endtoend_env_utils as eu
    assert os.path.realpath(e2e.__file__).startswith(os.path.realpath(args.reference))
    from env_build_b200 import synthetic as syn

    for task in ('left', 'straight', 'right'):
        rng = np.random.default_rng(syn.SEED_BASE + 77 + len(task))
        N, V = 400, eu.VEH_NUM[task]
        env = object.__new__(e2e.CrossroadEnd2end)
        env.dynamics = dm.VehicleDynamics()
        env.training_task = task
        env.num_future_data = 0
        env.env_model = dm.EnvironmentModel(task, 0)
        env.ego_l, env.ego_w = eu.L, eu.W
        env.ego_info_dim = 6
        traffic = object.__new__(tr.Traffic)
        rp = dm.ReferencePath(task, 0)
        paths = [tuple(np.asarray(a) for a in p) for p in rp.path_list]
        ref = syn.make_ref_indexes(rng, N)
        obs = syn.make_obs(rng, N, task, V, paths, ref)
        q = N // 8
        obs[q:6 * q, 9::4] = 400.0
        obs[2 * q:3 * q, 2], obs[2 * q:3 * q, 0] = 3.0, 9.0
        goal = dict(left=(-36.0, 5.6, 180.0), straight=(5.6, 36.0, 90.0), right=(36.0, -5.6, 0.0))[task]
        obs[3 * q:4 * q, 3], obs[3 * q:4 * q, 4], obs[3 * q:4 * q, 5] = goal
        obs[3 * q:4 * q, 1:3] = 0
        off = dict(left=(20.0, -40.0, 90.0), straight=(-10.0, -40.0, 90.0), right=(-3.0, -40.0, 90.0))[task]
        obs[4 * q:5 * q, 3], obs[4 * q:5 * q, 4], obs[4 * q:5 * q, 5] = off
        obs[5 * q:6 * q, 3] += 25.0
        obs[6 * q:7 * q, 6] = rng.choice([-20., 20., 14.9, 15.1], q)        # stored delta_y of the CURRENT obs
        act = syn.make_actions(rng, 1, N)[0]
        v_light = (rng.random(N) < 0.15).astype(np.int64)
        # the vehicles the done logic sees are those AFTER the traffic step; any set will do as input
        veh_after = syn.make_obs(rng, N, task, V, paths, ref)[:, 9:].reshape(N, V, 4)
        veh_after[q:6 * q] = 400.0
        veh_after[:, :, 0:2] = np.where(rng.random((N, V, 1)) < 0.04, obs[:, None, 3:5] + rng.uniform(-4, 4, (N, V, 2)),
                                        veh_after[:, :, 0:2])

        out = dict(obs=obs, act=act, v_light=v_light, veh_after=veh_after.astype(np.float32), ref=ref)
        scaled, nxt, par, corner, rbound, rew, code, coll = [], [], [], [], [], [], [], []
        parts = []
        for i in range(N):
            a = env._action_transformation_for_end2end(act[i])
            env.ego_dynamics = dict(v_x=obs[i, 0], v_y=obs[i, 1], r=obs[i, 2], x=obs[i, 3], y=obs[i, 4], phi=obs[i, 5])
            r, _info = env.compute_reward(obs[i], a)
            ns, npar = env._get_next_ego_state(a)
            dyn = env._get_ego_dynamics(ns, npar)
            # Traffic.collision_check on the ego just placed and the vehicles after the traffic step
            traffic.n_ego_dict = dict(ego=dict(x=dyn['x'], y=dyn['y'], phi=dyn['phi'], l=eu.L, w=eu.W))
            traffic.n_ego_vehicles = dict(ego=[dict(x=float(v[0]), y=float(v[1]), v=float(v[2]), phi=float(v[3]),
                                                    l=eu.L, w=eu.W) for v in veh_after[i]])
            traffic.collision_check()
            traffic.collision_flag = bool(traffic.n_ego_collision_flag['ego'])
            env.traffic = traffic
            env.ego_dynamics = dyn
            env.v_light = int(v_light[i])
            # _deviate_too_much reads delta_y from self.obs (the observation after the step)
            env.obs = np.concatenate([ns, obs[i, 6:9], veh_after[i].reshape(-1)]).astype(np.float32)
            typ, _d = env._judge_done()
            scaled.append(a); nxt.append(ns); par.append(npar)
            corner.append(np.array(dyn['Corner_point'], np.float64)); rbound.append(np.float64(dyn['r_bound']))
            rew.append(r); code.append(DONE.index(typ)); coll.append(traffic.collision_flag)
            parts.append([traffic.collision_flag, env._break_road_constrain(), env._deviate_too_much(),
                          env._break_stability(), env._break_red_light(), env._is_achieve_goal()])
        out.update(scaled=np.array(scaled, np.float32), next_ego=np.array(nxt, np.float32),
                   params=np.array(par, np.float32), corners=np.array(corner), r_bound=np.array(rbound),
                   reward=np.array(rew, np.float32), done_code=np.array(code, np.int8),
                   predicates=np.array(parts, bool),
                   _doc=np.array('unmodified reference endtoend.py:150-283,501-507 and traffic.py:263-295 on stand-in '
                                 'imports; rows = independent single-env calls'))
        # ---- _construct_veh_vector_short (E2E:340-464) on random scenes --------------------------
        classes = ('dl', 'du', 'dr', 'rd', 'rl', 'ru', 'ur', 'ud', 'ul', 'lu', 'lr', 'ld')
        route = dict(dl=('1o', '4i'), du=('1o', '3i'), dr=('1o', '2i'), rd=('2o', '1i'), rl=('2o', '4i'),
                     ru=('2o', '3i'), ur=('3o', '2i'), ud=('3o', '1i'), ul=('3o', '4i'), lu=('4o', '3i'),
                     lr=('4o', '2i'), ld=('4o', '1i'))
        S, NV = 300, 40
        sel_veh = np.zeros((S, NV, 4), np.float32)
        sel_veh[:, :, 0] = np.round(rng.uniform(-45, 45, (S, NV)) * 2) / 2         # half-metre grid: many key ties
        sel_veh[:, :, 1] = np.round(rng.uniform(-60, 50, (S, NV)) * 2) / 2
        sel_veh[:, :, 2] = rng.uniform(0, 8, (S, NV))
        sel_veh[:, :, 3] = rng.choice([0., 90., 180., -90.], (S, NV))
        sel_cls = rng.integers(-1, 12, (S, NV)).astype(np.int8)
        sel_cls[: S // 10] = -1                                                    # empty scenes: all fill values
        sel_ego = np.stack([rng.uniform(-30, 12, S), rng.uniform(-60, 30, S)], 1).astype(np.float32)
        sel_light = (rng.random(S) < 0.3).astype(np.int64)
        sel_virtual = rng.random(S) < 0.2
        sel_out = []
        for i in range(S):
            env.ego_dynamics = dict(x=sel_ego[i, 0], y=sel_ego[i, 1])
            env.v_light = int(sel_light[i])
            env.virtual_red_light_vehicle = bool(sel_virtual[i])
            env.all_vehicles = [dict(x=float(v[0]), y=float(v[1]), v=float(v[2]), phi=float(v[3]), l=4.8, w=2.0,
                                     route=route[classes[c]] if c >= 0 else ('9o', '9i'))
                                for v, c in zip(sel_veh[i], sel_cls[i])]
            sel_out.append(env._construct_veh_vector_short())
        out.update(sel_veh=sel_veh, sel_cls=sel_cls, sel_ego=sel_ego, sel_light=sel_light,
                   sel_virtual=sel_virtual, sel_out=np.array(sel_out, np.float32))
        # ---- _reset_init_state (E2E:472-499): the two np.random.random() draws replayed by seed ----
        K = 300
        rst_u, rst_ego, rst_path = np.zeros((K, 2)), np.zeros((K, 6), np.float32), np.zeros(K, np.int64)
        for i in range(K):
            k = i % 3
            env.ref_path = dm.ReferencePath(task, k)
            np.random.seed(1000 + i)
            rst_u[i] = np.random.random(), np.random.random()
            np.random.seed(1000 + i)
            ego = env._reset_init_state()['ego']
            rst_ego[i] = [ego['v_x'], ego['v_y'], ego['r'], float(ego['x']), float(ego['y']), float(ego['phi'])]
            rst_path[i] = k
        out.update(reset_u=rst_u, reset_ego=rst_ego, reset_path=rst_path)
        np.savez_compressed(os.path.join(HERE, 'env_%s.npz' % task), **out)
        print(task, 'done codes', np.bincount(out['done_code'], minlength=7))


if __name__ == '__main__':
    main()
