#!/usr/bin/env python
"""Model-based policy improvement through the differentiable model path: what the stop-gradients of
the reference's model (dynamics_and_models.py:195, 331, 402) are there for.  A small MLP policy is
rolled out closed loop for `horizon` steps of EnvironmentModel.rollout_out; the loss
  -mean_t rewards + penalty * mean_t punish_term_for_training
is differentiated through every step (ce2e_rollout_step_backward) back into the policy weights.

    python examples/train_policy.py              (needs a CUDA device and the built library)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from env_build_b200 import synthetic as syn                                   # noqa: E402
from env_build_b200.dynamics_and_models import EnvironmentModel               # noqa: E402
from env_build_b200.endtoend_env_utils import VEH_NUM                         # noqa: E402


def main(task='left', n_egos=4096, horizon=10, iters=60, penalty=10.0, seed=0):
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    model = EnvironmentModel(task, mode='training')
    V = VEH_NUM[task]
    ref = syn.make_ref_indexes(rng, n_egos)
    obs = torch.as_tensor(syn.make_obs(rng, n_egos, task, V, model.ref_path.path_list, ref), device='cuda')
    scale = obs.abs().mean(0).clamp_min(1.0)                     # crude input normalisation
    policy = torch.nn.Sequential(torch.nn.Linear(obs.shape[1], 64), torch.nn.Tanh(),
                                 torch.nn.Linear(64, 2), torch.nn.Tanh()).cuda()
    opt = torch.optim.Adam(policy.parameters(), lr=3e-3)
    first = last = None
    for it in range(iters):
        model.reset(obs, ref)
        cur, loss = obs, 0.
        for _ in range(horizon):
            actions = policy(cur / scale)
            cur, rewards, punish_train, _, _, _ = model.rollout_out(actions)
            loss = loss + (-rewards.mean() + penalty * punish_train.mean()) / horizon
        opt.zero_grad()
        loss.backward()
        opt.step()
        last = float(loss.detach())
        first = last if first is None else first
        if it % 10 == 0 or it == iters - 1:
            print('iter %3d  loss %.4f' % (it, last))
    print('loss %.4f -> %.4f over %d iterations (%d egos x %d steps each)' % (first, last, iters, n_egos, horizon))
    return first, last


if __name__ == '__main__':
    main()
