#!/usr/bin/env python
"""Where the e2e leg's time goes: each copy of one rollout alone (GB/s), then the pipelined leg."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from env_build_b200.dynamics_and_models import EnvironmentModel
from env_build_b200.rollout import RolloutGraph
dev = torch.device('cuda', 0)
B, H, V, D = 65536, 25, 32, 137
paths, obs, ref, tape = bench.make_inputs(B, 1)
model = EnvironmentModel('left', 0, mode='training', veh_mode_list=bench.mode_list())
h_obs = torch.from_numpy(obs).pin_memory(); h_tape = torch.from_numpy(tape).pin_memory(); h_ref = torch.from_numpy(ref).pin_memory()
h_out5 = torch.empty((H, 5, B)).pin_memory(); h_final = torch.empty((B, D)).pin_memory()
g = RolloutGraph(model, B, V, H)
g.load(h_obs, h_ref, h_tape); g.run(); torch.cuda.synchronize()


def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


dense = torch.empty((B, D), device=dev)
for name, fn, nbytes in (
        ('load(): H2D obs (dense host -> padded rows) + ref + tape', lambda: g.load(h_obs, h_ref, h_tape), obs.nbytes + ref.nbytes + tape.nbytes),
        ('H2D obs only into a dense device tensor', lambda: dense.copy_(h_obs, non_blocking=True), obs.nbytes),
        ('H2D tape only', lambda: g.tape.copy_(h_tape, non_blocking=True), tape.nbytes),
        ('D2H out5 (contiguous)', lambda: h_out5.copy_(g.out5, non_blocking=True), h_out5.numel() * 4),
        ('D2H final_obs (padded rows -> dense host)', lambda: h_final.copy_(g.final_obs, non_blocking=True), h_final.numel() * 4),
        ('D2H dense device tensor of the same size', lambda: h_final.copy_(dense, non_blocking=True), h_final.numel() * 4),
        ('graph run', g.run, 0)):
    ms = t(fn)
    print('%-62s %.3f ms %s' % (name, ms, ('%.1f GB/s' % (nbytes / ms / 1e6)) if nbytes else ''))
