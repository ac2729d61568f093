"""Time of ce2e_rollout_step_backward next to the forward step (CUDA events, L2-warm)."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from env_build_b200 import _lib, synthetic as syn
from env_build_b200.dynamics_and_models import EnvironmentModel, padded_rows

def t_us(fn, n=20):
    """n launches captured in one CUDA graph (the Python call costs more than a small kernel)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    for _ in range(3): g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / (5 * n)

lib = _lib.load()
for V in (32, 8):
    B, task = 65536, 'left'
    rng = np.random.default_rng(1)
    modes = syn.tiled_mode_list(['dl', 'dl', 'du', 'du', 'ud', 'ud', 'ul', 'ul'], V)
    model = EnvironmentModel(task, mode='training', veh_mode_list=modes)
    ref = syn.make_ref_indexes(rng, B)
    obs_h = syn.make_obs(rng, B, task, V, model.ref_path.path_list, ref)
    D = obs_h.shape[1]
    obs = padded_rows(B, D, 9, torch.device('cuda')); obs.copy_(torch.as_tensor(obs_h))
    act = torch.as_tensor(syn.make_actions(rng, 1, B)[0], device='cuda')
    dref = torch.as_tensor(ref, device='cuda', dtype=torch.int32)
    nxt = padded_rows(B, D, 9, torch.device('cuda')); out5 = torch.empty((5, B), device='cuda')
    g_next = torch.randn((B, 9), device='cuda'); g_out5 = torch.randn((5, B), device='cuda')
    g_obs = torch.zeros((B, 9), device='cuda'); g_act = torch.empty((B, 2), device='cuda')
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    h = model.ref_path.handle
    fwd = lambda: _lib.check(lib.ce2e_rollout_step(h, 0, vp(dref), vp(obs), obs.stride(0), vp(act), ctypes.byref(model._turn),
                                                   V, V, 0, vp(nxt), nxt.stride(0), vp(out5), None, B, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    bwd = lambda: _lib.check(lib.ce2e_rollout_step_backward(h, 0, vp(dref), vp(obs), obs.stride(0), vp(act), V, 0, vp(g_next), 9,
                                                            vp(g_out5), vp(g_obs), 9, vp(g_act), B, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    print('V=%d  forward %.1f us  backward %.1f us' % (V, t_us(fwd), t_us(bwd)))
    done = torch.empty((B,), dtype=torch.int8, device='cuda')
    sc = torch.zeros((B, 2), device='cuda')
    jd = lambda: _lib.check(lib.ce2e_judge_done(0, vp(nxt), nxt.stride(0), vp(sc), V, 0, 0, vp(done), B,
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    print('V=%d  judge_done %.1f us' % (V, t_us(jd)))
