"""How much of a k_model_step launch is fixed cost?  us per launch vs batch size and vehicle count
(CUDA-graph replay of 25 launches, CUDA events, L2 flushed between replays)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from env_build_b200 import synthetic as syn
from env_build_b200.dynamics_and_models import EnvironmentModel
from env_build_b200.endtoend_env_utils import VEHICLE_MODE_LIST
from env_build_b200.rollout import RolloutGraph

flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')


def us_per_launch(B, V, H=25):
    rng = np.random.default_rng(1)
    m = EnvironmentModel('left', 0, mode='training', veh_mode_list=syn.tiled_mode_list(VEHICLE_MODE_LIST['left'], V))
    ref = syn.make_ref_indexes(rng, B)
    g = RolloutGraph(m, B, V, H)
    g.load(syn.make_obs(rng, B, 'left', V, m.ref_path.path_list, ref), ref, syn.make_actions(rng, H, B))
    for _ in range(3):
        g.run()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    torch.cuda.synchronize()
    for a, b in ev:
        flush.zero_(); a.record(); g.run(); b.record()
    torch.cuda.synchronize()
    return 1e3 * sum(a.elapsed_time(b) for a, b in ev) / 10 / H


print('B      ' + ''.join('V=%-7d' % v for v in (0, 8, 16, 32)))
for B in (16, 2368, 4736, 16384, 32768, 65536, 131072):
    print('%-7d' % B + ''.join('%-9.2f' % us_per_launch(B, v) for v in (0, 8, 16, 32)))
