import cProfile, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from env_build_b200.dynamics_and_models import EnvironmentModel
dev = torch.device('cuda', 0)
B, H = 65536, 25
paths, obs, ref, tape = bench.make_inputs(B, 1)
model = EnvironmentModel('left', 0, mode='training', veh_mode_list=bench.mode_list())
d_tape = torch.from_numpy(tape).to(dev)
model.reset(obs, ref)
for k in range(H): model.rollout_out(d_tape[k])
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(100): model.rollout_out(d_tape[k % H])
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('100 calls: enqueue %.1f us/call, incl. sync %.1f us/call' % (1e4 * (t1 - t0), 1e4 * (t2 - t0)))
pr = cProfile.Profile(); pr.enable()
for k in range(200): model.rollout_out(d_tape[k % H])
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
