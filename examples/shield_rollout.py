#!/usr/bin/env python
"""The reference's online decision pattern (hierarchical_decision/hier_decision.py:89-129) on the
B200 path, for a batch of egos at once: evaluate every candidate path, then run the 5-step safety
shield rollout of all candidates in one batch.  The policy is a stand-in (random actions): the
reference's trained networks are not part of the model path.

    python examples/shield_rollout.py            (needs a CUDA device and the built library)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from env_build_b200 import synthetic as syn                                   # noqa: E402
from env_build_b200.dynamics_and_models import EnvironmentModel               # noqa: E402
from env_build_b200.endtoend_env_utils import VEH_NUM                         # noqa: E402


def main(task='left', n_egos=4096, horizon=5):
    rng = np.random.default_rng(0)
    model = EnvironmentModel(task, mode='training')
    obs = syn.make_obs(rng, n_egos, task, VEH_NUM[task], model.ref_path.path_list, 0)

    # one observation per (ego, candidate path): tracking columns re-projected per path
    cand, ref = model.candidate_observations(obs)                 # [3 * n_egos, D], [3 * n_egos]
    model.reset(cand, ref)
    unsafe = torch.zeros(cand.shape[0], device='cuda')
    ret = torch.zeros(cand.shape[0], device='cuda')
    for _ in range(horizon):                                      # is_safe, hier_decision.py:93-97
        actions = torch.rand((cand.shape[0], 2), device='cuda') * 2 - 1      # policy.run_batch(obs) stand-in
        obses, rewards, punish_train, punish_real, veh2veh4real, veh2road4real = model.rollout_out(actions)
        unsafe += veh2veh4real
        ret += rewards
    unsafe = unsafe.reshape(3, n_egos)
    ret = ret.reshape(3, n_egos)
    best = torch.where(unsafe > 0, torch.full_like(ret, -1e9), ret).argmax(0)    # best safe path per ego
    print('egos: %d, candidates unsafe within %d steps: %.1f %%, chosen path histogram: %s'
          % (n_egos, horizon, 100 * float((unsafe > 0).float().mean()), torch.bincount(best, minlength=3).tolist()))


if __name__ == '__main__':
    main()
