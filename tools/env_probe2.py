#!/usr/bin/env python
"""env step time under the bench's conditions (external zero action tensor) vs the probe's (uniform actions
written into env.action_buffer), for two builds."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from env_build_b200.endtoend import CrossroadEnd2end
B, V = 65536, 32
for kind in ('zeros-external', 'uniform-external', 'uniform-buffer', 'zeros-buffer'):
    env = CrossroadEnd2end('left', num_envs=B, veh_num=V, auto_reset=True, use_graph=True, reward_info=False)
    env.seed(1); env.reset()
    if kind.endswith('buffer'):
        act = env.action_buffer
    else:
        act = torch.zeros((B, 2), device='cuda')
    if kind.startswith('uniform'):
        act.uniform_(-1, 1)
    else:
        act.zero_()
    for _ in range(10): env.step(act)
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    e0 = int(env._bufs['episode'].sum())
    a.record()
    for _ in range(300): env.step(act)
    b.record(); torch.cuda.synchronize()
    print('%s: %.1f us per step; rows restarted per step %.0f' % (kind, a.elapsed_time(b) * 1e3 / 300, (int(env._bufs['episode'].sum()) - e0) / 300))
