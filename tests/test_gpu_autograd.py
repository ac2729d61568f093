"""Backward of EnvironmentModel.rollout_out (ce2e_rollout_step_backward through
env_build_b200/autograd.py) against torch.autograd on the float64 restatement in
oracle/torch_model.py (the same op graph TensorFlow differentiates in the reference).  Needs a GPU.
Tolerance: gradients are fp32 on the device and float64 in the oracle -> allclose(rtol=2e-3,
atol=2e-3); rows whose closest waypoint is a near-tie are excluded (different reference point)."""
import numpy as np
import pytest
import torch

from conftest import TASKS
from oracle import crossroad_oracle as orc
from oracle import torch_model as tm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dm():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from env_build_b200 import _lib
    if _lib.needs_build():
        _lib.build()
    from env_build_b200 import dynamics_and_models
    return dynamics_and_models


def _weights(rng, B, D9):
    return rng.normal(0, 1, (B, D9)), rng.normal(0, 1, (5, B))


def _close(a, b, what):
    bad = ~np.isclose(a, b, rtol=2e-3, atol=2e-3)
    assert bad.mean() < 2e-3, (what, int(bad.sum()), a[bad][:5], b[bad][:5])


@pytest.mark.parametrize('task', TASKS)
@pytest.mark.parametrize('steps', [1, 3])
def test_rollout_gradients(dm, task, steps):
    from env_build_b200 import synthetic as syn
    rng = np.random.default_rng(20210315 + steps)
    B, V = 1500, orc.VEH_NUM[task]
    model = dm.EnvironmentModel(task, mode='training')
    paths = model.ref_path.path_list
    ref = syn.make_ref_indexes(rng, B, out_of_range_frac=0.03)
    obs = syn.make_obs(rng, B, task, V, paths, ref)
    tape = syn.make_actions(rng, steps, B)
    w_obs, w_out = _weights(rng, B, 9)

    # device: leaf tensors on the GPU, loss = sum_t <w_out, out5_t> + <w_obs, next_obs_T[:, :9]>
    d_obs = torch.tensor(obs, device='cuda', requires_grad=True)
    d_act = [torch.tensor(tape[t], device='cuda', requires_grad=True) for t in range(steps)]
    model.reset(d_obs, ref)
    loss = 0.
    for t in range(steps):
        res = model.rollout_out(d_act[t])
        loss = loss + (torch.stack(res[1:]) * torch.tensor(w_out, device='cuda', dtype=torch.float32)).sum()
    loss = loss + (res[0][:, :9] * torch.tensor(w_obs, device='cuda', dtype=torch.float32)).sum()
    loss.backward()

    # oracle: float64 autograd on the CPU
    o_obs = torch.tensor(obs, dtype=torch.float64, requires_grad=True)
    o_act = [torch.tensor(tape[t], dtype=torch.float64, requires_grad=True) for t in range(steps)]
    cur, oloss = o_obs, 0.
    margin = np.full(B, np.inf)
    om = orc.EnvironmentModel(task, mode='training', path_list=paths)
    om.reset(obs, ref)
    for t in range(steps):
        _, mg = om.compute_next_obses(om.obses, orc.action_transformation(tape[t]), return_margin=True)
        margin = np.minimum(margin, mg)
        om.rollout_out(tape[t])
        res_o = tm.rollout_out(cur, o_act[t], task, ref, paths, orc.VEHICLE_MODE_LIST[task])
        oloss = oloss + (torch.stack(res_o[1:]) * torch.tensor(w_out)).sum()
        cur = res_o[0]
    oloss = oloss + (cur[:, :9] * torch.tensor(w_obs)).sum()
    oloss.backward()

    ok = margin > 1e-3
    assert ok.mean() > 0.9
    assert abs(float(loss.detach()) - float(oloss.detach())) <= 1e-3 * abs(float(oloss.detach())) + 1e-2 * steps or not ok.all()
    for t in range(steps):
        _close(d_act[t].grad.cpu().numpy()[ok], o_act[t].grad.numpy()[ok], 'd loss / d action[%d]' % t)
    g, go = d_obs.grad.cpu().numpy(), o_obs.grad.numpy()
    _close(g[ok][:, :9], go[ok][:, :9], 'd loss / d obs (ego + tracking)')
    assert np.abs(g[:, 9:]).max() == 0 and np.abs(go[:, 9:]).max() == 0        # stop_gradient on vehicles
    # clipped actions carry no gradient (tf.clip_by_value, DM:129)
    clipped = np.abs(tape[-1]) > 1.05
    assert np.abs(d_act[-1].grad.cpu().numpy()[clipped]).max(initial=0.0) == 0


def test_no_grad_path_unchanged(dm):
    """Without requires_grad the call takes the plain path and returns the same numbers."""
    from env_build_b200 import synthetic as syn
    rng = np.random.default_rng(4)
    task, B = 'left', 777
    model = dm.EnvironmentModel(task, mode='training')
    ref = syn.make_ref_indexes(rng, B)
    obs = syn.make_obs(rng, B, task, 8, model.ref_path.path_list, ref)
    act = syn.make_actions(rng, 1, B)[0]
    model.reset(obs, ref)
    plain = [r.numpy() for r in model.rollout_out(act)]
    model.reset(obs, ref)
    a = torch.tensor(act, device='cuda', requires_grad=True)
    diff = model.rollout_out(a)
    for p, d in zip(plain, diff):
        assert np.array_equal(p.view(np.int32), d.detach().cpu().numpy().view(np.int32))
    assert diff[1].requires_grad and diff[0].requires_grad


@pytest.mark.parametrize('V', [1, 5, 8, 32, 37])
def test_backward_tiled_equals_scalar(dm, V):
    """ce2e_rollout_step_backward has two kernels: rows with a 16 B aligned vehicle block take the tiled
    one (two lanes per row, vehicles staged through shared memory), anything else the scalar one.  Same
    gradients up to the summation order of the collision terms (two halves vs. one running sum)."""
    import ctypes
    from env_build_b200 import _lib, synthetic as syn
    rng = np.random.default_rng(100 + V)
    task, B = 'right', 1501                                     # ragged: last tile has 13 rows
    model = dm.EnvironmentModel(task, mode='training', veh_mode_list=syn.tiled_mode_list(orc.VEHICLE_MODE_LIST[task], V))
    ref = syn.make_ref_indexes(rng, B, out_of_range_frac=0.03)
    obs = syn.make_obs(rng, B, task, V, model.ref_path.path_list, ref)
    D = obs.shape[1]
    act = torch.as_tensor(syn.make_actions(rng, 1, B)[0], device='cuda')
    dref = torch.as_tensor(ref, device='cuda', dtype=torch.int32)
    g_next = torch.randn((B, 9), device='cuda')
    g_out5 = torch.randn((5, B), device='cuda')
    lib, vp = _lib.load(), (lambda t: ctypes.c_void_p(t.data_ptr()))
    got = []
    for padded in (True, False):
        if padded:
            o = dm.padded_rows(B, D, 9, torch.device('cuda'))
            o.copy_(torch.as_tensor(obs))
            assert (o.data_ptr() + 36) % 16 == 0 and o.stride(0) % 4 == 0
        else:
            o = torch.as_tensor(obs, device='cuda').contiguous()
            assert o.stride(0) % 4 != 0 or (o.data_ptr() + 36) % 16 != 0
        g_obs = torch.full((B, 9), float('nan'), device='cuda')
        g_act = torch.full((B, 2), float('nan'), device='cuda')
        _lib.check(lib.ce2e_rollout_step_backward(model.ref_path.handle, 0, vp(dref), vp(o), o.stride(0), vp(act), V, 0,
                                                  vp(g_next), 9, vp(g_out5), vp(g_obs), 9, vp(g_act), B, None))
        torch.cuda.synchronize()
        got.append((g_obs.cpu().numpy(), g_act.cpu().numpy()))
    for a, b in zip(got[0], got[1]):
        assert np.isfinite(a).all() and np.isfinite(b).all()
        assert np.allclose(a, b, rtol=1e-4, atol=1e-4), float(np.abs(a - b).max())


@pytest.mark.parametrize('task', TASKS)
def test_backward_against_reference_gradient_golden(dm, task):
    """ce2e_rollout_step_backward against tests/golden/grad_<task>.npz: the vector-Jacobian product of the
    UNMODIFIED reference rollout_out under the shim's GradientTape (TensorFlow's autodiff rules;
    make_golden_grad.py).  allclose(rtol=1e-4, atol=1e-4); rows whose closest waypoint is a near-tie
    (a different reference point in fp32) are excluded."""
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, 'grad_%s.npz' % task), allow_pickle=False))
    model = dm.EnvironmentModel(task, mode='training')
    obs = torch.tensor(g['obs'], device='cuda', requires_grad=True)
    act = torch.tensor(g['act'], device='cuda', requires_grad=True)
    model.reset(obs, g['ref'])
    res = model.rollout_out(act)
    # the forward of the differentiable path equals the reference's forward
    assert np.allclose(res[0].detach().cpu().numpy()[:, :6], g['next_obs'][:, :6], rtol=1e-5, atol=1e-5)
    for a, b in zip(res[1:], g['out5']):
        assert np.allclose(a.detach().cpu().numpy(), b, rtol=1e-5, atol=1e-5)
    loss = (res[0][:, :9] * torch.tensor(g['g_next9'], device='cuda')).sum() + \
        (torch.stack(res[1:]) * torch.tensor(g['g_out5'], device='cuda')).sum()
    loss.backward()
    om = orc.EnvironmentModel(task, mode='training', path_list=model.ref_path.path_list)
    om.reset(g['obs'], g['ref'])
    _, margin = om.compute_next_obses(g['obs'], orc.action_transformation(g['act']), return_margin=True)
    ok = margin > 1e-3
    assert ok.mean() > 0.95
    go, ga = obs.grad.cpu().numpy(), act.grad.cpu().numpy()
    assert np.allclose(go[ok][:, :9], g['grad_obs9'][ok], rtol=1e-4, atol=1e-4), \
        float(np.abs(go[ok][:, :9] - g['grad_obs9'][ok]).max())
    assert np.allclose(ga[ok], g['grad_act'][ok], rtol=1e-4, atol=1e-4), float(np.abs(ga[ok] - g['grad_act'][ok]).max())
    assert np.abs(go[:, 9:]).max() == 0
