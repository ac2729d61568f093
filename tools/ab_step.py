#!/usr/bin/env python
"""Same-box A/B of the fused step's two kernels: k_model_step_pair (TMA tensor maps, warp pair per 32
rows; ce2e_set_tma(1)) against k_model_step (cp.async, two lanes per row; ce2e_set_tma(0)).
CUDA-graph replay of H = 25 launches, CUDA events, L2 flushed between timed rollouts, variants
interleaved.  Optional argv: libraries to compare through CE2E_LIB-style paths are NOT handled here
(one process = one library); run once per build.

    python tools/ab_step.py > gpurun_out/ab_step.md      (on a B200)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from env_build_b200 import _lib, synthetic as syn                # noqa: E402
from env_build_b200.dynamics_and_models import EnvironmentModel  # noqa: E402
from env_build_b200.endtoend_env_utils import VEHICLE_MODE_LIST  # noqa: E402
from env_build_b200.rollout import RolloutGraph                  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
H = 25


def graph(task, B, V, tma):
    rng = np.random.default_rng(1)
    m = EnvironmentModel(task, 0, mode='training', veh_mode_list=syn.tiled_mode_list(VEHICLE_MODE_LIST[task], V))
    ref = syn.make_ref_indexes(rng, B)
    obs = syn.make_obs(rng, B, task, V, m.ref_path.path_list, ref)
    old = _lib.set_tma(tma)
    g = RolloutGraph(m, B, V, H)
    g.load(obs, ref, syn.make_actions(rng, H, B))
    g.run()                                   # captures with the current setting
    torch.cuda.synchronize()
    assert _lib.load().ce2e_last_step_kernel() == (2 if tma else 1)
    _lib.set_tma(old)
    return g


def main():
    reps = int(os.environ.get('AB_REPS', '15'))
    only_tma = os.environ.get('AB_TMA_ONLY') == '1'          # one column: compare builds across processes
    cfgs = [('left', 65536, 32), ('left', 131072, 32), ('left', 524288, 32), ('left', 65536, 8),
            ('straight', 65536, 9), ('right', 65536, 5), ('left', 4096, 32), ('left', 65536, 12), ('left', 65536, 16),
            ('left', 65536, 24), ('left', 524288, 8), ('straight', 524288, 9), ('right', 524288, 5), ('left', 262144, 8)]
    if os.environ.get('AB_CONFIGS'):
        cfgs = [cfgs[int(i)] for i in os.environ['AB_CONFIGS'].split(',')]
    print('# k_model_step_pair (TMA) vs k_model_step (cp.async), %s, lib %s\n' % (torch.cuda.get_device_name(0),
                                                                                 os.path.basename(_lib.LIB_PATH)))
    print('| config | TMA us/launch | frac | cp.async us/launch | frac | speed-up |\n|---|---|---|---|---|---|')
    for task, B, V in cfgs:
        gs = {t: graph(task, B, V, t) for t in ((True,) if only_tma else (True, False))}
        for g in gs.values():
            for _ in range(3):
                g.run()
        ts = {t: [] for t in gs}
        for _ in range(reps):
            for t in gs:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                flush.zero_()
                a.record()
                gs[t].run()
                b.record()
                torch.cuda.synchronize()
                ts[t].append(a.elapsed_time(b) * 1e3 / H)
        D = 9 + 4 * V
        med = {t: float(np.median(ts[t])) for t in ts}
        frac = {t: (8 * D + 32) * B / (med[t] * 1e-6) / 1e9 / PEAK for t in ts}
        if only_tma:
            print('| %s B=%d V=%d | %.2f | %.3f | | | |' % (task, B, V, med[True], frac[True]))
        else:
            print('| %s B=%d V=%d | %.2f | %.3f | %.2f | %.3f | %.3f |' % (task, B, V, med[True], frac[True], med[False],
                                                                          frac[False], med[False] / med[True]))
        del gs
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
