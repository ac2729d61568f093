#!/usr/bin/env python3
"""Golden GRADIENTS of EnvironmentModel.rollout_out, produced by the UNMODIFIED reference source.

    python tests/golden/make_golden_grad.py [--reference /root/reference]   (build container only)

The reference's `dynamics_and_models.py` is imported as it is on the NumPy-backed TensorFlow stand-in of
tests/golden/_refshim, whose `tf.GradientTape` implements reverse-mode autodiff with TensorFlow's
rules (tf.where -> selected branch, tf.clip_by_value -> closed interval, tf.argmin / tf.gather on
integer indices -> constant reference point, tf.stop_gradient on the vehicle columns, DM:195 / DM:331
/ DM:402).  One `rollout_out` step per task runs under the tape (mode='training', per-row paths,
native vehicle count); the vector-Jacobian product for random upstream gradients of the six outputs
is taken w.r.t. the observations and the normalised actions -- exactly what
ce2e_rollout_step_backward returns.  Output: tests/golden/grad_<task>.npz.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    args = ap.parse_args()
    sys.path.insert(0, os.path.join(HERE, '_refshim'))
    sys.path.insert(0, args.reference)
    sys.path.insert(0, ROOT)
    import dynamics_and_models as dm       # the reference's module
    import endtoend_env_utils as eu
    assert os.path.realpath(dm.__file__).startswith(os.path.realpath(args.reference)), dm.__file__
    from env_build_b200 import synthetic as syn
    tf = dm.tf

    for task in ('left', 'straight', 'right'):
        rng = np.random.default_rng(syn.SEED_BASE + 500 + len(task))
        B, V = 384, eu.VEH_NUM[task]
        model = dm.EnvironmentModel(task, 0, mode='training')
        paths = [tuple(np.asarray(a) for a in p) for p in model.ref_path.path_list]
        ref = syn.make_ref_indexes(rng, B, out_of_range_frac=0.03)
        obs = syn.make_obs(rng, B, task, V, paths, ref, near_frac=0.5)       # many active hinge terms
        act = syn.make_actions(rng, 1, B)[0]
        D = obs.shape[1]
        g_next = np.zeros((B, D), np.float32)
        g_next[:, :9] = rng.normal(0, 1, (B, 9))
        g_out = rng.normal(0, 1, (5, B)).astype(np.float32)

        t_obs, t_act = tf.constant(obs), tf.constant(act)
        with tf.GradientTape() as tape:
            tape.watch(t_obs)
            tape.watch(t_act)
            model.reset(t_obs, tf.constant(ref))
            res = model.rollout_out(t_act)
        grads = tape.gradient(list(res), [t_obs, t_act],
                              output_gradients=[g_next.astype(np.float64)] + [g.astype(np.float64) for g in g_out])
        g_obs, g_act = (np.asarray(g.numpy(), np.float64) for g in grads)
        assert np.isfinite(g_obs).all() and np.isfinite(g_act).all()
        assert np.abs(g_obs[:, 9:]).max() == 0.0, 'vehicle columns carry tf.stop_gradient'
        out = dict(obs=obs, ref=ref, act=act, g_next9=g_next[:, :9].copy(), g_out5=g_out,
                   grad_obs9=g_obs[:, :9].copy(), grad_act=g_act,
                   next_obs=np.asarray(res[0].numpy()), out5=np.stack([np.asarray(r.numpy()) for r in res[1:]]),
                   _doc=np.array('vector-Jacobian product of the unmodified reference rollout_out (DM:118-126) under the '
                                 'shim GradientTape; upstream gradients g_next9 / g_out5, float64 accumulation'))
        np.savez_compressed(os.path.join(HERE, 'grad_%s.npz' % task), **out)
        print(task, 'grad_obs9 rms %.3g, grad_act rms %.3g, nonzero hinge rows %d' % (
            np.sqrt((g_obs ** 2).mean()), np.sqrt((g_act ** 2).mean()), int((out['out5'][1] > 0).sum())))


if __name__ == '__main__':
    main()
