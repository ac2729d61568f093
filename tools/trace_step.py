#!/usr/bin/env python
"""Per-warp phase timeline of ONE k_model_step_pair launch (debug build with -DCE2E_TRACE, selected
through CE2E_LIB).  Stamps (clock64 of the warp's SM, first tile only): 0 entry, 1 after the prologue,
2 after griddepcontrol.wait, 3 ego columns landed + sincos, 4 ego phase done, 5+2c / 6+2c chunk c
landed / stored, 13 flush done, 14 tile done.  Prints, per role, the median / p90 / max of each
phase's duration and of the absolute end time relative to the SM's first stamp.

    nvcc ... -DCE2E_TRACE -o build/libce2e_trace.so ...
    CE2E_LIB=build/libce2e_trace.so python tools/trace_step.py [B] [V]
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from env_build_b200 import _lib, synthetic as syn                # noqa: E402
from env_build_b200.dynamics_and_models import EnvironmentModel, _ptr, padded_rows  # noqa: E402
from env_build_b200.endtoend_env_utils import VEHICLE_MODE_LIST  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
V = int(sys.argv[2]) if len(sys.argv) > 2 else 32
task = 'left'
rng = np.random.default_rng(1)
m = EnvironmentModel(task, 0, mode='training', veh_mode_list=syn.tiled_mode_list(VEHICLE_MODE_LIST[task], V))
ref = torch.as_tensor(syn.make_ref_indexes(rng, B), device='cuda')
obs_h = syn.make_obs(rng, B, task, V, m.ref_path.path_list, ref.cpu().numpy())
D = 9 + 4 * V
a = padded_rows(B, D, 9)
b = padded_rows(B, D, 9)
a.copy_(torch.as_tensor(obs_h, device='cuda'))
act = torch.as_tensor(syn.make_actions(rng, 1, B)[0], device='cuda')
out5 = torch.empty((5, B), device='cuda')
NW = 296 * 14          # = 148 * 28
trace = torch.zeros((NW, 16), dtype=torch.int64, device='cuda')
lib = _lib.load()
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
# ce2e_env_step takes a dict16 pointer: the trace build writes its stamps there
done = torch.empty((B,), dtype=torch.int8, device='cuda')
scaled = torch.empty((B, 2), device='cuda')
WARM = os.environ.get('TRACE_WARM') == '1'      # input just written by the previous launch (as inside a rollout)
for rep in range(4):
    if not WARM:
        flush.zero_()
    trace.zero_()
    _lib.check(lib.ce2e_env_step(m.ref_path.handle, _ptr(ref), _ptr(a), a.stride(0), _ptr(act), ctypes.byref(m._turn), V, 0,
                                 0, _ptr(b), b.stride(0), _ptr(out5), _ptr(trace), _ptr(scaled), _ptr(done), B, None))
    a, b = b, a
    torch.cuda.synchronize()
t = trace.cpu().numpy()
used = t[:, 14] != 0
t = t[used]
sm = t[:, 15]
W = int(os.environ.get('TRACE_WARPS', '28'))         # warps per block of the traced build
w = np.nonzero(used)[0] % W
role = ((w ^ (w >> 2)) & 1) if os.environ.get('TRACE_PLAIN_ROLES') != '1' else (w & 1)
names = {1: 'prologue', 2: 'griddep wait', 3: 'ego loads+sincos', 4: 'ego phase', 5: 'chunk0 wait', 6: 'chunk0', 7: 'chunk1 wait',
         8: 'chunk1', 9: 'chunk2 wait', 10: 'chunk2', 11: 'chunk3 wait', 12: 'chunk3', 13: 'flush', 14: 'exchange+out'}
t0 = np.zeros_like(t[:, 0])
for s_ in np.unique(sm):
    t0[sm == s_] = t[sm == s_, 0].min()
print('# phase timeline (%s), B=%d V=%d, %d warps traced, clock cycles (1.965 GHz)\n' % ('warm L2' if WARM else 'L2 flushed', B, V, len(t)))
for r in (0, 1):
    print('## role %d (%s)\n\n| phase | dur median | p90 | max | end min | end p10 | end median | end p90 | end max |\n|---|---|---|---|---|---|---|---|---|' %
          (r, 'reward warp, first half' if r == 0 else 'dynamics warp, second half'))
    sel = role == r
    prev = 0
    for k in range(1, 15):
        if not (t[sel, k] != 0).all():
            continue
        d = t[sel, k] - t[sel, prev]
        e = t[sel, k] - t0[sel]
        print('| %s | %d | %d | %d | %d | %d | %d | %d | %d |' % (names[k], np.median(d), np.percentile(d, 90), d.max(), e.min(),
                                                                np.percentile(e, 10), np.median(e), np.percentile(e, 90), e.max()))
        prev = k
    print()
span = np.array([t[sm == s_, 14].max() - t[sm == s_, 0].min() for s_ in np.unique(sm)])
print('per-SM span (first entry to last tile end): median %d, max %d cycles' % (np.median(span), span.max()))
