import sys, os, torch
sys.path.insert(0, '.')
from env_build_b200.endtoend import CrossroadEnd2end
env = CrossroadEnd2end('left', num_envs=65536, veh_num=32, auto_reset=True, use_graph=False, reward_info=False)
env.seed(1); env.reset(); act = env.action_buffer; act.uniform_(-1, 1)
for _ in range(30): env.step(act)
torch.cuda.synchronize()
