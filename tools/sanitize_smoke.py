"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck): a few fused steps with
ragged batch sizes, both modes, several vehicle counts, padded and unpadded rows, plus the env step."""
import numpy as np
import torch

from env_build_b200 import synthetic as syn
from env_build_b200.dynamics_and_models import EnvironmentModel
from env_build_b200.endtoend import CrossroadEnd2end
from env_build_b200.endtoend_env_utils import VEHICLE_MODE_LIST

rng = np.random.default_rng(0)
for task in ('left', 'right'):
    for V in (0, 5, 9, 32, 37):
        for mode in ('selecting', 'training'):
            m = EnvironmentModel(task, mode=mode, veh_mode_list=syn.tiled_mode_list(VEHICLE_MODE_LIST[task], V))
            B = 1000 + V
            ref = syn.make_ref_indexes(rng, B, out_of_range_frac=0.05)
            obs = syn.make_obs(rng, B, task, V, m.ref_path.path_list, ref if mode == 'training' else 1)
            if mode == 'training':
                m.reset(obs, ref)
            else:
                m.add_traj(obs, 1)
            for _ in range(2):
                res = m.rollout_out(syn.make_actions(rng, 1, B)[0])
            m.compute_rewards(obs, np.zeros((B, 2), np.float32))
            m.ss(obs, np.zeros((B, 2), np.float32))
env = CrossroadEnd2end('straight', num_envs=777, auto_reset=True)
env.reset()
for _ in range(3):
    env.step(rng.uniform(-1, 1, (777, 2)).astype(np.float32))
torch.cuda.synchronize()
print('sanitize smoke done')
