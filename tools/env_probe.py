#!/usr/bin/env python
"""us per CrossroadEnd2end.step() of the batched, auto-resetting environment (graph replay vs eager)."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from env_build_b200.endtoend import CrossroadEnd2end
B, V = 65536, 32
for graph in (True, False):
    env = CrossroadEnd2end('left', num_envs=B, veh_num=V, auto_reset=True, use_graph=graph, reward_info=False)
    env.seed(1); env.reset()
    act = env.action_buffer
    act.uniform_(-1, 1)
    for _ in range(10): env.step(act)
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(500): env.step(act)
    b.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print('graph=%s: %.1f us per step (CUDA events), %.1f us wall; done rows so far %d' % (
        graph, a.elapsed_time(b) * 1e3 / 500, (t1 - t0) * 1e6 / 500, int(env._bufs['episode'].sum()) - B))
