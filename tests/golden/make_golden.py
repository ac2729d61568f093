#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference source.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py [--reference /root/reference]

TensorFlow, bezier and matplotlib are absent from this image; the reference
modules `dynamics_and_models` and `endtoend_env_utils` are imported from the
reference tree AS THEY ARE, on top of the NumPy-backed stand-ins in
tests/golden/_refshim/ (fp32 element-wise ops with TF's scalar-conversion
rules; sin/cos/atan via float64; `bezier.Curve.evaluate_multi` restated from
the package's published algorithm).  So the expression trees, operator
precedence, loop orders and branch structure that produced these vectors are
the reference's own; only the element-wise op backend is substituted.

Every array written here is an input or an output of a reference function; the
file names its source (file:line) in `_doc`.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def _load_reference(ref_dir):
    sys.path.insert(0, os.path.join(HERE, '_refshim'))
    sys.path.insert(0, ref_dir)
    sys.path.insert(0, ROOT)
    import dynamics_and_models as dm   # noqa: E402  (the reference's module)
    import endtoend_env_utils as eu    # noqa: E402
    assert os.path.realpath(dm.__file__).startswith(os.path.realpath(ref_dir)), dm.__file__
    return dm, eu


def _np(x):
    return np.asarray(x.numpy() if hasattr(x, 'numpy') else x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    args = ap.parse_args()
    dm, eu = _load_reference(args.reference)
    from env_build_b200 import synthetic as syn

    tf = dm.tf
    out = {}
    # ------------------------------------------------------------------ constants
    const = dict(L=eu.L, W=eu.W, LANE_WIDTH=eu.LANE_WIDTH, LANE_NUMBER=eu.LANE_NUMBER,
                 CROSSROAD_SIZE=eu.CROSSROAD_SIZE, EXPECTED_V=eu.EXPECTED_V)
    common = {'const_' + k: np.float64(v) for k, v in const.items()}
    for task in ('left', 'straight', 'right'):
        common['mode_list_' + task] = np.array(eu.VEHICLE_MODE_LIST[task])
    vd = dm.VehicleDynamics()
    for k, v in vd.vehicle_params.items():
        common['vp_' + k] = np.float64(v)

    # ------------------------------------------------------------------ f_xu (DM:52-87)
    rng = np.random.default_rng(syn.SEED_BASE)
    B = 256
    st = np.stack([rng.uniform(0, 12, B), rng.uniform(-1, 1, B), rng.uniform(-0.6, 0.6, B),
                   rng.uniform(-60, 60, B), rng.uniform(-60, 60, B), rng.uniform(-200, 400, B)], 1).astype(np.float32)
    st[:8, 0] = 0.0                                   # standstill rows
    st[8] = [5, 0, 0, 0, 0, 90]                       # SURVEY 8c hand-derivable KAT
    st[9] = [8, 0.2, 0.1, 1.875, -30, 92]             # SURVEY 8c scratch KAT
    ac = np.stack([rng.uniform(-0.42, 0.42, B), rng.uniform(-3.2, 1.7, B)], 1).astype(np.float32)
    ac[8] = [0, 0]
    ac[9] = [0.1, -1.5]
    nxt, par = vd.f_xu(tf.constant(st), tf.constant(ac), 0.1)
    nxt2, par2 = vd.prediction(st, ac, 10)            # NumPy inputs, the Gym path (E2E:277-279)
    common.update(fxu_states=st, fxu_actions=ac, fxu_next=_np(nxt), fxu_params=_np(par),
                  pred_next=_np(nxt2), pred_params=_np(par2))

    # ------------------------------------------------------------------ action scaling (DM:128-132)
    act_norm = np.concatenate([rng.uniform(-1.5, 1.5, (60, 2)),
                               [[1, 1], [-1, -1], [2, 0], [0, -2]]]).astype(np.float32)
    model_l = dm.EnvironmentModel('left', 0, mode='selecting')
    common.update(act_norm=act_norm, act_scaled=_np(model_l._action_transformation_for_end2end(tf.constant(act_norm))))

    # ------------------------------------------------------------------ phi wrap (DM:577-580)
    pd = np.concatenate([rng.uniform(-720, 720, 100), [180, -180, 180.00002, -180.00002, 540, -540]]).astype(np.float32)
    common.update(phidiff_in=pd, phidiff_out=_np(dm.deal_with_phi_diff(tf.constant(pd))))

    # ------------------------------------------------------------------ predict_for_a_mode (DM:405-427)
    Bv = 128
    vehs = np.stack([rng.uniform(-40, 40, Bv), rng.uniform(-40, 40, Bv), rng.uniform(0, 10, Bv),
                     rng.uniform(-180, 180, Bv)], 1).astype(np.float32)
    vehs[:6, 3] = [180, -180, 179.9, -179.9, 90, -90]
    vehs[6:10, 0] = [25, -25, 24.999, -24.999]
    common['pfm_in'] = vehs
    for mode in ('dl', 'rd', 'ur', 'lu', 'dr', 'ru', 'ul', 'ld', 'du', 'ud', 'lr', 'rl'):
        common['pfm_out_' + mode] = _np(model_l.predict_for_a_mode(tf.constant(vehs), mode))

    # ------------------------------------------------------------------ judge_feasible / deal_with_phi (EU)
    jf_xy = rng.uniform(-70, 70, (400, 2))
    jf_xy[:12] = [[0, -25], [3.75, -25], [1, -25], [1, -25.0001], [-25, 1], [25, -1], [-25.0001, 1],
                  [25.0001, -1], [5, 25], [5, 25.0001], [0, 0], [11.25, 30]]
    common['jf_xy'] = jf_xy
    for task in ('left', 'straight', 'right'):
        common['jf_' + task] = np.array([eu.judge_feasible(x, y, task) for x, y in jf_xy])
    dwp = np.array([0, 180, -180, 181, -181, 540, -540, 359.5, 720.25])
    common.update(dwp_in=dwp, dwp_out=np.array([eu.deal_with_phi(float(p)) for p in dwp]))
    common['_doc'] = np.array('reference outputs: dynamics_and_models.py f_xu/prediction (52-87), '
                              '_action_transformation_for_end2end (128-132), deal_with_phi_diff (577-580), '
                              'predict_for_a_mode (405-427); endtoend_env_utils.py judge_feasible (73-104), '
                              'deal_with_phi (232-237)')
    np.savez_compressed(os.path.join(HERE, 'common.npz'), **common)

    # ------------------------------------------------------------------ per task
    for ti, task in enumerate(('left', 'straight', 'right')):
        g = {}
        rp = dm.ReferencePath(task, 0)
        for pi_, p in enumerate(rp.path_list):
            g['path%d_x' % pi_], g['path%d_y' % pi_], g['path%d_phi' % pi_] = (np.asarray(a) for a in p)
        g['path_len_list'] = np.array(rp.path_len_list)
        g['control_points'] = np.array(rp.control_points, dtype=np.float64)
        paths = [tuple(np.asarray(a) for a in p) for p in rp.path_list]
        Vn = eu.VEH_NUM[task]
        rng = np.random.default_rng(syn.SEED_BASE + 10 + ti)

        # -- tracking_error_vector / find_closest_point / future_n_data (DM:702-770)
        Bt = 96
        for pi_ in range(3):
            rp.set_path(pi_)
            ob = syn.make_obs(rng, Bt, task, 0, paths, pi_)
            xs, ys, phis, vs = ob[:, 3].copy(), ob[:, 4].copy(), ob[:, 5].copy(), ob[:, 0].copy()
            # the reference's own script inputs (DM:805-808) + far-away / seam poses
            xs[:4], ys[:4] = [1.875, 1.875, -10, -20], [-20, 0, -10, -1]
            phis[:4], vs[:4] = [90, 135, 135, 180], [10, 12, 10, 10]
            xs[4:8], ys[4:8] = [80, -80, 0, 30], [80, -80, -90, 30]
            phis[4:8] = [-179, 179, 270, -270]
            # exactly on decimated waypoints -> zero lateral error
            xs[8:12] = paths[pi_][0][[0, 1000, 2000, 3000]]
            ys[8:12] = paths[pi_][1][[0, 1000, 2000, 3000]]
            phis[8:12] = paths[pi_][2][[0, 1000, 2000, 3000]]
            vs[8:12] = 8.0
            g['trk%d_in' % pi_] = np.stack([xs, ys, phis, vs], 1)
            for n in (0, 3, 10):
                g['trk%d_n%d' % (pi_, n)] = _np(rp.tracking_error_vector(xs, ys, phis, vs, n))
            idx, pts = rp.find_closest_point(xs, ys)
            g['fcp%d_idx' % pi_] = _np(idx)
            g['fcp%d_pts' % pi_] = np.stack([_np(p) for p in pts], 1)
            idx5, pts5 = rp.find_closest_point(xs, ys, ratio=5)
            g['fcp%d_idx_r5' % pi_] = _np(idx5)
            fut = rp.future_n_data(tf.constant(np.array([600, 0, 3500, len(paths[pi_][0]) - 3], np.int64)), 5)
            g['fut%d' % pi_] = np.stack([np.stack([_np(c) for c in f], 1) for f in fut], 0)   # [5, 4, 3]

        # -- compute_rewards at native V and at V=32 (DM:186-320)
        for V in (Vn, 32):
            Br = 128
            ob = syn.make_obs(rng, Br, task, V, paths, syn.make_ref_indexes(rng, Br))
            # a pad-vehicle row (E2E:440-447 far-away fill values) -> veh2veh exactly 0
            ob[0, 9:] = np.tile(np.array([1.875, -55, 0, 90], np.float32), V)
            ob[0, 3:6] = [1.875, -30, 90]
            an = syn.make_actions(rng, 1, Br)[0]
            model = dm.EnvironmentModel(task, 0, mode='selecting')
            asc = model._action_transformation_for_end2end(tf.constant(an))
            r = model.compute_rewards(tf.constant(ob), asc)
            g['rew_V%d_obs' % V] = ob
            g['rew_V%d_act' % V] = an
            g['rew_V%d_out5' % V] = np.stack([_np(t) for t in r[:5]], 1)
            g['rew_V%d_dict' % V] = np.stack([_np(r[5][k]) for k in sorted(r[5].keys())], 1)
            g['rew_dict_keys'] = np.array(sorted(r[5].keys()))

        # -- rollout_out, H=25 (DM:118-126): BASELINE config #1 (B=1) and small batches
        H = 25
        native_list = list(eu.VEHICLE_MODE_LIST[task])
        for tag, V, Bm, mode, n in (('cfg1', Vn, 1, 'selecting', 0),
                                    ('selV', Vn, 12, 'selecting', 0),
                                    ('trnV', Vn, 12, 'training', 0),
                                    ('sel32', 32, 8, 'selecting', 0),
                                    ('trn32', 32, 8, 'training', 0),
                                    ('seln10', Vn, 6, 'selecting', 10),
                                    ('trnn3', Vn, 6, 'training', 3)):
            # V=32 needs a length-32 mode list: a DATA patch of the reference's table
            # (SURVEY.md section 0 item 3); restored right after.
            eu.VEHICLE_MODE_LIST[task] = syn.tiled_mode_list(native_list, V)
            dm.VEHICLE_MODE_LIST[task] = eu.VEHICLE_MODE_LIST[task]
            try:
                model = dm.EnvironmentModel(task, n, mode=mode)
                path_index = (ti + len(tag)) % 3
                if mode == 'training':
                    ref = syn.make_ref_indexes(rng, Bm, out_of_range_frac=0.15)
                    ref[0] = 3
                else:
                    ref = np.full(Bm, path_index, np.int32)
                ob0 = syn.make_obs(rng, Bm, task, V, paths, ref, num_future_data=n)
                tape = syn.make_actions(rng, H, Bm)
                if mode == 'training':
                    model.reset(tf.constant(ob0), tf.constant(ref))
                else:
                    model.add_traj(tf.constant(ob0), path_index)
                obs_seq, out_seq = [], []
                for t in range(H):
                    res = model.rollout_out(tf.constant(tape[t]))
                    obs_seq.append(_np(res[0]))
                    out_seq.append(np.stack([_np(x) for x in res[1:]], 1))
                g['ro_%s_obs0' % tag] = ob0
                g['ro_%s_ref' % tag] = ref
                g['ro_%s_path' % tag] = np.int32(path_index)
                g['ro_%s_tape' % tag] = tape
                g['ro_%s_obs' % tag] = np.stack(obs_seq, 0)      # [H, B, D]
                g['ro_%s_out5' % tag] = np.stack(out_seq, 0)     # [H, B, 5]
                if tag in ('selV', 'sel32'):
                    # barrier shield (DM:134-184) on the same first-step inputs
                    model.add_traj(tf.constant(ob0), path_index)
                    g['ss_%s' % tag] = _np(model.ss(tf.constant(ob0), tf.constant(tape[0]), lam=0.1))
            finally:
                eu.VEHICLE_MODE_LIST[task] = native_list
                dm.VEHICLE_MODE_LIST[task] = native_list
        g['_doc'] = np.array('reference outputs for task %s: ReferencePath tables (DM:598-700), '
                             'tracking_error_vector/find_closest_point/future_n_data (DM:702-770), '
                             'compute_rewards (DM:186-320), rollout_out x25 (DM:118-126), ss (DM:134-184)' % task)
        np.savez_compressed(os.path.join(HERE, 'task_%s.npz' % task), **g)
        print(task, 'written;', len(g), 'arrays')


if __name__ == '__main__':
    main()
