#!/usr/bin/env python
"""Per-config numbers for the BASELINE.json config list (single-GPU shares), written as a markdown
table.  Not the driver's bench (bench.py is); same timing hygiene: warm-up, CUDA events, L2 flush
between timed iterations, CUDA-graph replay for the H-step rollouts.

    python tools/bench_configs.py > profiles/r01_configs.md      (on a B200)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from env_build_b200 import synthetic as syn                      # noqa: E402
from env_build_b200.dynamics_and_models import EnvironmentModel, VehicleDynamics   # noqa: E402
from env_build_b200.endtoend import CrossroadEnd2end             # noqa: E402
from env_build_b200.endtoend_env_utils import VEHICLE_MODE_LIST, VEH_NUM          # noqa: E402
from env_build_b200.rollout import RolloutGraph                  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if \
    os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / n


def rollout(task, B, V, mode, H=25, n=0):
    rng = np.random.default_rng(1)
    m = EnvironmentModel(task, n, mode=mode, veh_mode_list=syn.tiled_mode_list(VEHICLE_MODE_LIST[task], V))
    ref = syn.make_ref_indexes(rng, B)
    obs = syn.make_obs(rng, B, task, V, m.ref_path.path_list, ref if mode == 'training' else 0, n)
    g = RolloutGraph(m, B, V, H)
    if mode != 'training':
        m.ref_path.set_path(0)
    g.load(obs, ref, syn.make_actions(rng, H, B))
    ms = timed(g.run)
    D = 6 + 3 * (n + 1) + 4 * V
    us = 1e3 * ms / H
    return B * H / (ms * 1e-3), us, (8 * D + 32) * B / (us * 1e-6) / 1e9 / PEAK


rows = []
# config 2: batch=4096 ego-only dynamics step
rng = np.random.default_rng(0)
st = torch.tensor(np.stack([rng.uniform(0, 12, 4096), rng.uniform(-1, 1, 4096), rng.uniform(-.5, .5, 4096),
                            rng.uniform(-60, 60, 4096), rng.uniform(-60, 60, 4096), rng.uniform(-180, 180, 4096)], 1),
                  dtype=torch.float32, device='cuda')
ac = torch.tensor(rng.uniform(-0.4, 0.4, (4096, 2)), dtype=torch.float32, device='cuda')
vd = VehicleDynamics()
g = torch.cuda.CUDAGraph()
vd.f_xu(st, ac, 0.1)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    for _ in range(100):
        vd.f_xu(st, ac, 0.1)
ms = timed(g.replay)
rows.append(('#2 f_xu (k_dynamics_step), B=4096, 100 launches per graph', '%.2f us per launch' % (10 * ms),
             '%.3g rows/s' % (4096 * 100 / (ms * 1e-3)), 'launch-latency bound: 0.29 MB per launch'))
for label, task, B, V, mode in [
        ('#3 rollout_out H=25, B=65536, V=32, training (headline)', 'left', 65536, 32, 'training'),
        ('#3 same, mode=selecting (one path for all rows)', 'left', 65536, 32, 'selecting'),
        ('#3 native V=8 (task left, D=41)', 'left', 65536, 8, 'training'),
        ('#3 native V=9 (task straight, D=45)', 'straight', 65536, 9, 'training'),
        ('#3 native V=5 (task right, D=29)', 'right', 65536, 5, 'training'),
        ('#4 per-GPU share: B=65536, 3 ref paths (training), V=32', 'left', 65536, 32, 'training'),
        ('#5 per-GPU share: B=131072, V=32', 'left', 131072, 32, 'training'),
        ('large batch B=524288, V=32 (HBM bound)', 'left', 524288, 32, 'training')]:
    v, us, frac = rollout(task, B, V, mode)
    rows.append((label, '%.2f us per launch' % us, '%.3g env-steps/s' % v, 'roofline frac %.3f' % frac))
# f-1: batched env step (two launches per step)
env = CrossroadEnd2end('left', num_envs=65536, veh_num=32)
env.reset()
act = torch.zeros((65536, 2), device='cuda')
ms = timed(lambda: env.step(act), n=10)
rows.append(('f-1 CrossroadEnd2end.step, 65536 envs, V=32 (fused step + done kernel, Python call included)',
             '%.1f us per step' % (1e3 * ms), '%.3g env-steps/s' % (65536 / (ms * 1e-3)), ''))
print('# Per-config measurements (one B200, %s)\n' % torch.cuda.get_device_name(0))
print('| config | time | throughput | note |\n|---|---|---|---|')
for r in rows:
    print('| %s | %s | %s | %s |' % r)
print('\nroofline frac = (8 D + 32) B bytes per launch / measured %.1f GB/s.' % PEAK)
