import time, numpy as np, torch, sys
sys.path.insert(0, '.')
import bench
from env_build_b200.dynamics_and_models import EnvironmentModel
from env_build_b200.rollout import RolloutGraph
dev = torch.device('cuda', 0)
B, H, V = 65536, 25, 32
paths, obs, ref, tape = bench.make_inputs(B, 1)
model = EnvironmentModel('left', 0, mode='training', veh_mode_list=bench.mode_list())
h_obs = torch.from_numpy(obs).pin_memory(); h_tape = torch.from_numpy(tape).pin_memory(); h_ref = torch.from_numpy(ref).pin_memory()
d = torch.empty_like(h_obs, device=dev); h_back = torch.empty_like(h_obs).pin_memory()
def t(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
print('H2D 36MB  %.3f ms -> %.1f GB/s' % (1e3 * t(lambda: d.copy_(h_obs, non_blocking=True)), obs.nbytes / t(lambda: d.copy_(h_obs, non_blocking=True)) / 1e9))
print('D2H 36MB  %.3f ms -> %.1f GB/s' % (1e3 * t(lambda: h_back.copy_(d, non_blocking=True)), obs.nbytes / t(lambda: h_back.copy_(d, non_blocking=True)) / 1e9))
s2 = torch.cuda.Stream()
def both():
    d.copy_(h_obs, non_blocking=True)
    with torch.cuda.stream(s2): h_back.copy_(d, non_blocking=True)
print('H2D+D2H concurrently %.3f ms' % (1e3 * t(both)))
d_tape = torch.from_numpy(tape).to(dev)
model.reset(d, torch.from_numpy(ref).to(dev))
def api25():
    for k in range(H): model.rollout_out(d_tape[k])
print('25 x rollout_out, device-resident inputs: %.3f ms (kernel time ~0.47 ms)' % (1e3 * t(api25)))
g = RolloutGraph(model, B, V, H)
h_out5 = torch.empty((H, 5, B)).pin_memory(); h_final = torch.empty((B, 137)).pin_memory()
def graph_e2e():
    g.load(h_obs, h_ref, h_tape); g.run()
    h_out5.copy_(g.out5, non_blocking=True); h_final.copy_(g.final_obs, non_blocking=True)
graph_e2e()
print('RolloutGraph e2e (H2D obs+tape, graph, D2H out5+final): %.3f ms' % (1e3 * t(graph_e2e)))
