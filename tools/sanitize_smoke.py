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
# the TMA pair kernel's other work splits (balanced: chosen for many tiles per pair; forced here), several tiles per pair
from env_build_b200 import _lib
for mode_tma, B in ((3, 2100), (2, 2100), (1, 70000)):
    old = _lib.set_tma(mode_tma)
    m = EnvironmentModel('left', mode='training', veh_mode_list=syn.tiled_mode_list(VEHICLE_MODE_LIST['left'], 32))
    ref = syn.make_ref_indexes(rng, B, out_of_range_frac=0.05)
    m.reset(syn.make_obs(rng, B, 'left', 32, m.ref_path.path_list, ref), ref)
    for _ in range(2):
        m.rollout_out(syn.make_actions(rng, 1, B)[0])
    _lib.set_tma(old)
env = CrossroadEnd2end('straight', num_envs=777, auto_reset=True)
env.reset()
for _ in range(3):
    env.step(rng.uniform(-1, 1, (777, 2)).astype(np.float32))
# backward pass, vehicle selection, done logic, horizon-fused mode, candidate paths
from env_build_b200.rollout import RolloutGraph
m = EnvironmentModel('left', mode='training')
B = 1003
ref = syn.make_ref_indexes(rng, B)
obs = torch.tensor(syn.make_obs(rng, B, 'left', 8, m.ref_path.path_list, ref), device='cuda', requires_grad=True)
act = torch.tensor(syn.make_actions(rng, 1, B)[0], device='cuda', requires_grad=True)
m.reset(obs, ref)
res = m.rollout_out(act)
(res[1].sum() + res[2].sum() + res[0][:, :9].sum()).backward()
env1 = CrossroadEnd2end('left', num_envs=333)
env1.reset()
env1.construct_veh_vectors(rng.uniform(-40, 40, (333, 17, 4)).astype(np.float32), rng.integers(-1, 12, (333, 17)),
                           rng.uniform(-30, 10, (333, 2)).astype(np.float32), rng.integers(0, 2, 333))
env1._judge_done()
for V in (0, 5, 32):
    mm = EnvironmentModel('right', mode='training', veh_mode_list=syn.tiled_mode_list(VEHICLE_MODE_LIST['right'], V))
    g = RolloutGraph(mm, 517, V, 4, use_graph=False, fused=True)
    rf = syn.make_ref_indexes(rng, 517)
    g.load(syn.make_obs(rng, 517, 'right', V, mm.ref_path.path_list, rf), rf, syn.make_actions(rng, 4, 517))
    g.run()
m2 = EnvironmentModel('straight', mode='training')
m2.candidate_observations(syn.make_obs(rng, 129, 'straight', 9, m2.ref_path.path_list, 0))
torch.cuda.synchronize()
print('sanitize smoke done')
