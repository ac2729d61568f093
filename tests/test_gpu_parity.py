"""CUDA path (libce2e.so through env_build_b200's reference-shaped API) against the oracle and
the golden vectors generated from the unmodified reference.  Needs a GPU.

Tolerances: integer / index results and every function without a transcendental
(find_closest_point, tracking_error_vector, action scaling, indexs2points) must be BIT-EXACT;
functions through sin/cos/atan (f_xu, compute_rewards, veh_predict, rollout_out) must satisfy
allclose(rtol=1e-5, atol=1e-5) -- the tolerance BASELINE.json's north_star states.
Rows whose nearest-waypoint decision is a near-tie in the oracle (margin between the best and
second-best squared distance < 1e-4 m^2) are excluded from next-obs tracking comparisons: a
1-ulp difference in sin/cos can legitimately move such a row to the neighbouring waypoint."""
import os

import numpy as np
import pytest
import torch

from conftest import TASKS, golden_paths
from oracle import crossroad_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = ATOL = 1e-5


@pytest.fixture(scope='module')
def dm():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from env_build_b200 import _lib
    if _lib.needs_build():
        _lib.build()
    from env_build_b200 import dynamics_and_models
    return dynamics_and_models


def bits_equal(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype == np.float32, (a.shape, b.shape, a.dtype, b.dtype)
    ok = (a.view(np.int32) == b.view(np.int32)) | ((a == 0) & (b == 0)) | (np.isnan(a) & np.isnan(b))
    assert ok.all(), (int((~ok).sum()), a[~ok][:5], b[~ok][:5])


def close(a, b, what=''):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = ~np.isclose(a, b, rtol=RTOL, atol=ATOL, equal_nan=True)
    assert not bad.any(), (what, int(bad.sum()), a[bad][:5], b[bad][:5], float(np.abs(a - b)[bad].max()))


def tiled(task, V):
    from env_build_b200.synthetic import tiled_mode_list
    return tiled_mode_list(orc.VEHICLE_MODE_LIST[task], V)


# ------------------------------------------------------------------------------------------
# VehicleDynamics
# ------------------------------------------------------------------------------------------
def test_f_xu_golden(dm, golden_common):
    c = golden_common
    vd = dm.VehicleDynamics()
    nxt, par = vd.f_xu(c['fxu_states'], c['fxu_actions'], 0.1)
    close(nxt.numpy(), c['fxu_next'], 'f_xu next')
    close(par.numpy(), c['fxu_params'], 'f_xu params')
    nxt, par = vd.prediction(c['fxu_states'], c['fxu_actions'], 10)
    close(nxt.numpy(), c['pred_next'], 'prediction')
    # everything except x', y' (the only columns through sin/cos) is plain IEEE arithmetic
    bits_equal(nxt.numpy()[:, [0, 1, 2, 5]], c['pred_next'][:, [0, 1, 2, 5]])
    assert vd.vehicle_params['F_zf'] == float(c['vp_F_zf']) and vd.vehicle_params['F_zr'] == float(c['vp_F_zr'])


def test_dynamics_config2(dm):
    """BASELINE config #2: batch=4096 ego-only dynamics step."""
    rng = np.random.default_rng(20210312)
    B = 4096
    st = np.stack([rng.uniform(0, 12, B), rng.uniform(-1, 1, B), rng.uniform(-0.5, 0.5, B), rng.uniform(-60, 60, B),
                   rng.uniform(-60, 60, B), rng.uniform(-180, 180, B)], 1).astype(np.float32)
    st[:16, 0] = 0.0                                   # standstill rows (no singularity, DM:76/78)
    ac = orc.action_transformation(rng.uniform(-1.2, 1.2, (B, 2)).astype(np.float32))
    want, wpar = orc.f_xu(st, ac, 0.1)
    got, gpar = dm.VehicleDynamics().f_xu(st, ac, 0.1)
    close(got.numpy(), want)
    bits_equal(got.numpy()[:, [0, 1, 2, 5]], want[:, [0, 1, 2, 5]])
    ok = st[:, 0] > 0.05                               # alpha = atan(./(v_x+1e-8)) is ill-conditioned at rest
    close(gpar.numpy()[ok], wpar[ok])
    close(gpar.numpy()[:, 2:], wpar[:, 2:])
    # hand KAT: heading 90 deg, no action
    n, _ = dm.VehicleDynamics().f_xu(np.array([[5, 0, 0, 0, 0, 90]], np.float32), np.zeros((1, 2), np.float32), 0.1)
    n = n.numpy()[0]
    assert n[0] == 5 and n[1] == 0 and n[2] == 0 and n[4] == 0.5 and n[5] == 90 and abs(n[3] + 2.1856e-8) < 1e-9


def test_action_scaling(dm, golden_common):
    m = dm.EnvironmentModel('left')
    bits_equal(m._action_transformation_for_end2end(golden_common['act_norm']).numpy(), golden_common['act_scaled'])


def test_phi_wrap(dm, golden_common):
    bits_equal(dm.deal_with_phi_diff(golden_common['phidiff_in']).numpy(), golden_common['phidiff_out'])


# ------------------------------------------------------------------------------------------
# ReferencePath
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('task', TASKS)
def test_tracking_golden(dm, task, golden_task):
    g = golden_task(task)
    rp = dm.ReferencePath(task, 0, path_list=golden_paths(g))          # the reference's own tables
    for pi in range(3):
        rp.set_path(pi)
        xs, ys, phis, vs = (np.ascontiguousarray(c) for c in g['trk%d_in' % pi].T)
        for n in (0, 3, 10):
            bits_equal(rp.tracking_error_vector(xs, ys, phis, vs, n).numpy(), g['trk%d_n%d' % (pi, n)])
        idx, pts = rp.find_closest_point(xs, ys)
        assert idx.dtype == torch.int64 and (idx.numpy() == g['fcp%d_idx' % pi]).all()
        bits_equal(np.stack([p.numpy() for p in pts], 1), g['fcp%d_pts' % pi])
        idx5, _ = rp.find_closest_point(xs, ys, ratio=5)
        assert (idx5.numpy() == g['fcp%d_idx_r5' % pi]).all()
        fut = rp.future_n_data(np.array([600, 0, 3500, len(rp.path[0]) - 3], np.int64), 5)
        bits_equal(np.stack([np.stack([c.numpy() for c in f], 1) for f in fut], 0), g['fut%d' % pi])
        pts2 = rp.indexs2points(np.array([-5, 0, 17, 10 ** 6], np.int64))
        want = orc.ReferencePath(task, pi, path_list=golden_paths(g)).indexs2points(np.array([-5, 0, 17, 10 ** 6]))
        bits_equal(np.stack([p.numpy() for p in pts2]), np.stack(want))


@pytest.mark.parametrize('task', TASKS)
def test_tracking_random_bit_exact(dm, task):
    """Own tables, 20k random poses over the whole map (incl. far outside), per-row paths."""
    rng = np.random.default_rng(5)
    B = 20000
    rp = dm.ReferencePath(task, 0)
    xs = rng.uniform(-80, 80, B).astype(np.float32)
    ys = rng.uniform(-80, 80, B).astype(np.float32)
    phis = rng.uniform(-360, 360, B).astype(np.float32)
    vs = rng.uniform(0, 35, B).astype(np.float32)
    ref = rng.integers(-1, 5, B).astype(np.int32)
    got = rp.tracking_error_vector(xs, ys, phis, vs, 4, ref_indexes=ref).numpy()
    want = np.zeros_like(got)
    for pi in range(3):
        o = orc.ReferencePath(task, pi, path_list=rp.path_list)
        m = ref == pi
        want[m] = o.tracking_error_vector(xs[m], ys[m], phis[m], vs[m], 4)
    bits_equal(got, want)
    assert np.abs(got[(ref < 0) | (ref > 2)]).max() == 0


@pytest.mark.parametrize('task', TASKS)
def test_candidate_grid_equals_brute_force_on_device(dm, task):
    """4M query points per path: the grid-restricted scan (what the fused step runs) returns the
    same index as scanning all candidates, bit for bit -- including points on cell edges, on
    waypoints, on bisectors between waypoints, far off the map, inf and NaN."""
    rng = np.random.default_rng(77)
    rp = dm.ReferencePath(task, 0)
    n = 1 << 20
    for pi in range(3):
        rp.set_path(pi)
        wx, wy = rp.path[0][::10], rp.path[1][::10]
        k = rng.integers(0, len(wx), n)
        k1 = np.minimum(k + 1, len(wx) - 1)
        sets = [np.stack([wx[k] + rng.normal(0, 2.0, n), wy[k] + rng.normal(0, 2.0, n)], 1),
                np.stack([rng.uniform(-90, 90, n), rng.uniform(-90, 90, n)], 1),
                np.stack([np.round(rng.uniform(-80, 80, n) * 2) / 2, wy[k] + rng.normal(0, 1.0, n)], 1),
                np.stack([(wx[k] + wx[k1]) / 2 + rng.normal(0, 1e-6, n), (wy[k] + wy[k1]) / 2], 1)]
        pts = np.concatenate(sets).astype(np.float32)
        pts[:8] = [[np.nan, 0], [0, np.nan], [np.inf, 0], [0, -np.inf], [1e30, 1e30], [-1e6, 3], [0, 0], [wx[5], wy[5]]]
        xs, ys = torch.as_tensor(pts[:, 0], device='cuda'), torch.as_tensor(pts[:, 1], device='cuda')
        fast, fpts = rp.find_closest_point(xs, ys)
        slow, spts = rp.find_closest_point(xs, ys, brute_force=True)
        assert torch.equal(fast, slow)
        for a, b in zip(fpts, spts):
            assert torch.equal(a.view(torch.int32), b.view(torch.int32))


# ------------------------------------------------------------------------------------------
# EnvironmentModel pieces
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('task', TASKS)
def test_compute_rewards_golden(dm, task, golden_task):
    g = golden_task(task)
    keys = list(g['rew_dict_keys'])
    for V in (orc.VEH_NUM[task], 32):
        m = dm.EnvironmentModel(task, veh_mode_list=tiled(task, V))
        ob, an = g['rew_V%d_obs' % V], g['rew_V%d_act' % V]
        res = m.compute_rewards(ob, m._action_transformation_for_end2end(an))
        close(np.stack([r.numpy() for r in res[:5]], 1), g['rew_V%d_out5' % V], 'out5 V=%d' % V)
        close(np.stack([res[5][k].numpy() for k in keys], 1), g['rew_V%d_dict' % V], 'dict V=%d' % V)
        assert sorted(res[5].keys()) == keys
        want = g['rew_V%d_out5' % V]
        # zero stays exactly zero (callers test `punish > 0`, hier_decision.py:97)
        got = np.stack([r.numpy() for r in res[:5]], 1)
        assert ((got[:, 1:] == 0) == (want[:, 1:] == 0)).all()


def test_predict_for_a_mode_golden(dm, golden_common):
    c = golden_common
    m = dm.EnvironmentModel('left')
    for mode in ('dl', 'rd', 'ur', 'lu', 'dr', 'ru', 'ul', 'ld', 'du', 'ud', 'lr', 'rl'):
        got = m.predict_for_a_mode(c['pfm_in'], mode).numpy()
        close(got, c['pfm_out_' + mode], mode)
        bits_equal(got[:, 2:], c['pfm_out_' + mode][:, 2:])        # v and phi: no sin/cos involved


@pytest.mark.parametrize('task', TASKS)
def test_veh_and_ego_predict(dm, task):
    rng = np.random.default_rng(11)
    from env_build_b200 import synthetic as syn
    m = dm.EnvironmentModel(task)
    V = orc.VEH_NUM[task]
    obs = syn.make_obs(rng, 777, task, V, m.ref_path.path_list, 0)
    close(m.veh_predict(obs[:, 9:]).numpy(), orc.veh_predict(obs[:, 9:], orc.VEHICLE_MODE_LIST[task]))
    act = orc.action_transformation(syn.make_actions(rng, 1, 777)[0])
    close(m.ego_predict(obs[:, :6], act).numpy(), orc.ego_predict(obs[:, :6], act))


# ------------------------------------------------------------------------------------------
# rollout_out
# ------------------------------------------------------------------------------------------
ROLLOUT_TAGS = ['cfg1', 'selV', 'trnV', 'sel32', 'trn32', 'seln10', 'trnn3']


def _models(dm, g, task, tag):
    ob0 = g['ro_%s_obs0' % tag]
    n = {'seln10': 10, 'trnn3': 3}.get(tag, 0)
    V = (ob0.shape[1] - 6 - 3 * (n + 1)) // 4
    mode = 'training' if tag.startswith('trn') else 'selecting'
    model = dm.EnvironmentModel(task, n, mode=mode, veh_mode_list=tiled(task, V))
    model.ref_path = dm.ReferencePath(task, 0, path_list=golden_paths(g))
    om = orc.EnvironmentModel(task, n, mode=mode, path_list=golden_paths(g), veh_mode_list=tiled(task, V))
    return model, om, mode, n


@pytest.mark.parametrize('task', TASKS)
@pytest.mark.parametrize('tag', ROLLOUT_TAGS)
def test_rollout_teacher_forced(dm, task, tag, golden_task):
    """Every step of the golden H=25 rollouts, re-seeded from the reference's own obs_t."""
    g = golden_task(task)
    model, om, mode, n = _models(dm, g, task, tag)
    ref, tape = g['ro_%s_ref' % tag], g['ro_%s_tape' % tag]
    ntr = 3 * (n + 1)
    prev = g['ro_%s_obs0' % tag]
    for t in range(tape.shape[0]):
        if mode == 'training':
            model.reset(prev, ref)
            om.reset(prev, ref)
        else:
            model.add_traj(prev, int(g['ro_%s_path' % tag]))
            om.add_traj(prev, int(g['ro_%s_path' % tag]))
        res = model.rollout_out(tape[t])
        want_obs, want5 = g['ro_%s_obs' % tag][t], g['ro_%s_out5' % tag][t]
        _, margin = om.compute_next_obses(prev, orc.action_transformation(tape[t]), return_margin=True)
        got = res[0].numpy()
        close(np.stack([r.numpy() for r in res[1:]], 1), want5, '%s out5 t=%d' % (tag, t))
        close(got[:, :6], want_obs[:, :6], 'ego')
        close(got[:, 6 + ntr:], want_obs[:, 6 + ntr:], 'veh')
        ok = margin > 1e-4
        close(got[ok, 6:6 + ntr], want_obs[ok, 6:6 + ntr], 'tracking')
        close(model.actions.numpy(), orc.action_transformation(tape[t]))
        prev = want_obs
    if mode == 'training':
        assert ref[0] == 3 and np.abs(got[0, 6:6 + ntr]).max() == 0     # unmatched ref_index -> zeros


@pytest.mark.parametrize('task', TASKS)
@pytest.mark.parametrize('tag', ['selV', 'trn32'])
def test_rollout_free_running(dm, task, tag, golden_task):
    """Free-running H=25 (state fed back on the device): drift stays far below the 0.33 m
    waypoint granularity unless a near-tie flips; reports the max error."""
    g = golden_task(task)
    model, om, mode, n = _models(dm, g, task, tag)
    ref, tape = g['ro_%s_ref' % tag], g['ro_%s_tape' % tag]
    if mode == 'training':
        model.reset(g['ro_%s_obs0' % tag], ref)
    else:
        model.add_traj(g['ro_%s_obs0' % tag], int(g['ro_%s_path' % tag]))
    worst = 0.0
    for t in range(tape.shape[0]):
        res = model.rollout_out(tape[t])
        want = g['ro_%s_obs' % tag][t]
        got = res[0].numpy()
        worst = max(worst, float(np.abs(got[:, :6] - want[:, :6]).max()))
        assert np.allclose(got[:, :6], want[:, :6], rtol=1e-4, atol=1e-4)
        assert np.allclose(got[:, 9 + 3 * n:], want[:, 9 + 3 * n:], rtol=1e-4, atol=1e-4)
    print('free-running max |ego err| over 25 steps: %.3g' % worst)


@pytest.mark.parametrize('task', TASKS)
def test_ss_golden(dm, task, golden_task):
    g = golden_task(task)
    for tag in ('selV', 'sel32'):
        model, om, mode, n = _models(dm, g, task, tag)
        model.add_traj(g['ro_%s_obs0' % tag], int(g['ro_%s_path' % tag]))
        got = model.ss(g['ro_%s_obs0' % tag], g['ro_%s_tape' % tag][0], lam=0.1).numpy()
        close(got, g['ss_%s' % tag], 'ss')


@pytest.mark.parametrize('task', TASKS)
@pytest.mark.parametrize('mode', ['selecting', 'training'])
@pytest.mark.parametrize('V', [0, 5, 9, 32, 40])
def test_rollout_synthetic(dm, task, mode, V):
    """Seeded synthetic batch (ragged size, edge rows, near vehicles) against the oracle."""
    from env_build_b200 import synthetic as syn
    rng = np.random.default_rng(20210313 + V)
    B = 3001
    model = dm.EnvironmentModel(task, mode=mode, veh_mode_list=tiled(task, V))
    ref = syn.make_ref_indexes(rng, B, out_of_range_frac=0.03)
    obs = syn.make_obs(rng, B, task, V, model.ref_path.path_list, ref if mode == 'training' else 2)
    act = syn.make_actions(rng, 1, B)[0]
    om = orc.EnvironmentModel(task, mode=mode, path_list=model.ref_path.path_list, veh_mode_list=tiled(task, V))
    if mode == 'training':
        model.reset(obs, ref)
        om.reset(obs, ref)
    else:
        model.add_traj(obs, 2)
        om.add_traj(obs, 2)
    res = model.rollout_out(act)
    sc = orc.action_transformation(act)
    want5 = orc.compute_rewards(obs, sc, task)[:5]
    want, margin = om.compute_next_obses(obs, sc, return_margin=True)
    got = res[0].numpy()
    for a, b in zip(res[1:], want5):
        close(a.numpy(), b)
    close(got[:, :6], want[:, :6])
    close(got[:, 9:], want[:, 9:])
    ok = margin > 1e-4
    assert ok.mean() > 0.99
    close(got[ok, 6:9], want[ok, 6:9])
    # the separately callable halves agree bit-for-bit with the fused step
    r5 = model.compute_rewards(obs, sc)[:5]
    for a, b in zip(res[1:], r5):
        bits_equal(a.numpy(), b.numpy())
    if mode == 'training':
        model.reset(obs, ref)
    else:
        model.add_traj(obs, 2)
    bits_equal(model.compute_next_obses(obs, sc).numpy(), got)


def test_rollout_drops_unlisted_vehicles(dm):
    """DM:398-402: only len(VEHICLE_MODE_LIST[task]) vehicles are predicted, rewards use all."""
    from env_build_b200 import synthetic as syn
    rng = np.random.default_rng(3)
    task, B = 'right', 257
    model = dm.EnvironmentModel(task, mode='selecting')                 # native list: 5 vehicles
    obs = syn.make_obs(rng, B, task, 8, model.ref_path.path_list, 0)
    act = syn.make_actions(rng, 1, B)[0]
    model.add_traj(obs, 0)
    res = model.rollout_out(act)
    om = orc.EnvironmentModel(task, mode='selecting', path_list=model.ref_path.path_list)
    om.add_traj(obs, 0)
    want = om.rollout_out(act)
    assert res[0].shape == (B, 29) == want[0].shape
    close(res[0].numpy()[:, :6], want[0][:, :6])
    close(res[0].numpy()[:, 9:], want[0][:, 9:])
    for a, b in zip(res[1:], want[1:]):
        close(a.numpy(), b)


# ------------------------------------------------------------------------------------------
# size-independent properties at BASELINE sizes
# ------------------------------------------------------------------------------------------
def test_full_size_properties(dm):
    """B=65536, V=32 (config #3): results do not depend on how the batch is split (the kernel picks
    a different lanes-per-row factor for small batches), nor on row padding; far vehicles give
    exactly zero collision penalty."""
    from env_build_b200 import synthetic as syn
    rng = np.random.default_rng(20210313)
    task, B, V = 'left', 65536, 32
    model = dm.EnvironmentModel(task, mode='training', veh_mode_list=tiled(task, V))
    ref = syn.make_ref_indexes(rng, B, out_of_range_frac=0.01)
    obs = syn.make_obs(rng, B, task, V, model.ref_path.path_list, ref)
    act = syn.make_actions(rng, 1, B)[0]
    model.reset(obs, ref)
    full = [r.numpy() for r in model.rollout_out(act)]
    # (a) shard invariance: 3 ragged shards, concatenated
    cuts = [0, 1000, 30001, B]
    parts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        model.reset(obs[a:b], ref[a:b])
        parts.append([r.numpy() for r in model.rollout_out(act[a:b])])
    for i in range(6):
        bits_equal(np.concatenate([p[i] for p in parts]), full[i])
    # (b) unpadded rows straight through the C ABI (scalar vehicle loads) == padded rows
    from env_build_b200 import _lib
    import ctypes
    dev_obs = torch.as_tensor(obs[:5001], device='cuda').contiguous()
    assert dev_obs.stride(0) == 137
    out = torch.empty((5001, 137), device='cuda')
    out5 = torch.empty((5, 5001), device='cuda')
    dact = torch.as_tensor(act[:5001], device='cuda')
    dref = torch.as_tensor(ref[:5001], device='cuda')
    _lib.check(_lib.load().ce2e_rollout_step(model.ref_path.handle, 0, dref.data_ptr(), dev_obs.data_ptr(), 137,
                                             dact.data_ptr(), ctypes.byref(model._turn), V, V, 0, out.data_ptr(), 137,
                                             out5.data_ptr(), None, 5001, None))
    torch.cuda.synchronize()
    model.reset(obs[:5001], ref[:5001])
    pad = [r.numpy() for r in model.rollout_out(act[:5001])]
    bits_equal(out.cpu().numpy(), pad[0])
    bits_equal(out5.cpu().numpy(), np.stack(pad[1:]))
    # (c) all vehicles far away -> veh2veh exactly 0
    far = obs.copy()
    far[:, 9::4] = 500.0
    model.reset(far, ref)
    res = model.rollout_out(act)
    assert float(res[4].abs().max()) == 0.0
    # (d) statistical sanity against the oracle on a 4096-row sample
    sel = rng.choice(B, 4096, replace=False)
    want5 = orc.compute_rewards(obs[sel], orc.action_transformation(act[sel]), task)[:5]
    for a, b in zip(full[1:], want5):
        close(a[sel], b)


def test_offsets_beyond_4GiB(dm):
    """Maximum sizes: 8 Mi rows x 144 floats = 4.8 GB per buffer, so byte offsets pass 2^31 and 2^32.
    The big batch is a small one tiled on the device; every tile must reproduce the small batch's
    results bit for bit (rows are independent, the kernel's summation order is fixed)."""
    from env_build_b200 import synthetic as syn
    from env_build_b200.dynamics_and_models import padded_rows
    rng = np.random.default_rng(77)
    task, V, Bs, reps = 'left', 32, 8192, 1024
    model = dm.EnvironmentModel(task, mode='training', veh_mode_list=tiled(task, V))
    ref = syn.make_ref_indexes(rng, Bs, out_of_range_frac=0.01)
    obs = syn.make_obs(rng, Bs, task, V, model.ref_path.path_list, ref)
    act = syn.make_actions(rng, 1, Bs)[0]
    model.reset(obs, ref)
    small = [r.clone() for r in model.rollout_out(act)]
    B = Bs * reps
    big = padded_rows(B, 137, 9, torch.device('cuda'))
    assert big.stride(0) * 4 * B > (1 << 32)
    big.unflatten(0, (reps, Bs)).copy_(torch.as_tensor(obs, device='cuda').unsqueeze(0).expand(reps, Bs, 137))
    model.reset(big, torch.as_tensor(ref, device='cuda').repeat(reps))
    del big
    res = model.rollout_out(torch.as_tensor(act, device='cuda').repeat(reps, 1))
    torch.cuda.synchronize()
    assert res[0].shape == (B, 137)
    for k in (0, 1, reps // 2 - 1, reps // 2, reps - 2, reps - 1):      # tiles on both sides of 2^31 and 2^32 B
        assert torch.equal(res[0][k * Bs:(k + 1) * Bs], small[0]), k
    for i in range(1, 6):
        assert torch.equal(res[i].view(reps, Bs), small[i].unsqueeze(0).expand(reps, Bs)), i
    del res
    torch.cuda.empty_cache()


def test_edge_cases_and_errors(dm):
    m = dm.EnvironmentModel('straight', mode='selecting')
    m.ref_path.set_path(0)
    # empty batch
    m.add_traj(np.zeros((0, 45), np.float32), 0)
    res = m.rollout_out(np.zeros((0, 2), np.float32))
    assert res[0].shape == (0, 45) and res[1].shape == (0,)
    # batch of one (reference callers: hier_decision.py:101)
    from env_build_b200 import synthetic as syn
    rng = np.random.default_rng(1)
    obs = syn.make_obs(rng, 1, 'straight', 9, m.ref_path.path_list, 0)
    m.add_traj(obs, 0)
    res = m.rollout_out(np.array([[0.5, 0.0]], np.float32))
    om = orc.EnvironmentModel('straight', mode='selecting', path_list=m.ref_path.path_list)
    om.add_traj(obs, 0)
    want = om.rollout_out(np.array([[0.5, 0.0]], np.float32))
    close(res[0].numpy(), want[0])
    assert len(res[0]) == 1 and res[1].numpy()[0] == pytest.approx(float(want[1][0]), rel=1e-5)
    # error behaviour
    with pytest.raises(AssertionError):
        dm.ReferencePath('u-turn')
    with pytest.raises(ValueError):
        m.add_traj(np.zeros((4, 44), np.float32), 0)
    with pytest.raises(ValueError):
        m.add_traj(obs, 0)
        m.rollout_out(np.zeros((2, 2), np.float32))
    with pytest.raises(IndexError):
        m.add_traj(obs, 7)
    mt = dm.EnvironmentModel('left', mode='training')
    mt.reset(np.zeros((2, 41), np.float32))
    with pytest.raises(ValueError):
        mt.rollout_out(np.zeros((2, 2), np.float32))
    # NaN propagates, nothing traps
    bad = obs.copy()
    bad[0, 3] = np.nan
    m.add_traj(bad, 0)
    r = m.rollout_out(np.array([[0.0, 0.0]], np.float32))
    assert np.isnan(r[0].numpy()[0, 3])


def test_fast_trig_option_within_tolerance(dm):
    """ce2e_set_fast_trig: vehicle sin/cos from the special-function unit stays inside 1e-5."""
    from env_build_b200 import _lib
    from env_build_b200 import synthetic as syn
    rng = np.random.default_rng(9)
    task, B, V = 'left', 20000, 32
    model = dm.EnvironmentModel(task, mode='training', veh_mode_list=tiled(task, V))
    ref = syn.make_ref_indexes(rng, B)
    obs = syn.make_obs(rng, B, task, V, model.ref_path.path_list, ref)
    act = syn.make_actions(rng, 1, B)[0]
    model.reset(obs, ref)
    exact = [r.numpy() for r in model.rollout_out(act)]
    assert _lib.set_fast_trig(True) is False
    try:
        model.reset(obs, ref)
        fast = [r.numpy() for r in model.rollout_out(act)]
    finally:
        assert _lib.set_fast_trig(False) is True
    sc = orc.action_transformation(act)
    want5 = orc.compute_rewards(obs, sc, task)[:5]
    for a, b in zip(fast[1:], want5):
        close(a, b)
    om = orc.EnvironmentModel(task, mode='training', path_list=model.ref_path.path_list, veh_mode_list=tiled(task, V))
    om.reset(obs, ref)
    close(fast[0][:, 9:], om.compute_next_obses(obs, sc)[:, 9:])
    bits_equal(fast[0][:, :9], exact[0][:, :9])                 # the ego path is untouched
    print('fast-trig max |dx| vs exact kernel: %.3g' % np.abs(fast[0] - exact[0]).max())


@pytest.mark.parametrize('task', TASKS)
def test_candidate_paths_shield_rollout(dm, task):
    """The reference's decision pattern (hier_decision.py:89-119): one observation per candidate path,
    then a 5-step shield rollout of every candidate -- here all 3B candidates in one batch."""
    from env_build_b200 import synthetic as syn
    rng = np.random.default_rng(21)
    B, V = 500, orc.VEH_NUM[task]
    model = dm.EnvironmentModel(task, mode='training')
    paths = model.ref_path.path_list
    obs = syn.make_obs(rng, B, task, V, paths, 0)
    cand, ref = model.candidate_observations(obs)
    assert cand.shape == (3 * B, obs.shape[1]) and ref.numpy().tolist() == [0] * B + [1] * B + [2] * B
    c = cand.numpy()
    for p in range(3):                                            # env.set_traj(path); env._get_obs()
        rp = orc.ReferencePath(task, p, path_list=paths)
        want = rp.tracking_error_vector(obs[:, 3], obs[:, 4], obs[:, 5], obs[:, 0], 0)
        bits_equal(c[p * B:(p + 1) * B, 6:9], want)
        bits_equal(c[p * B:(p + 1) * B, :6], obs[:, :6])
        bits_equal(c[p * B:(p + 1) * B, 9:], obs[:, 9:])
    model.reset(cand, ref)
    prev = c
    unsafe = torch.zeros(3 * B, device='cuda')
    unsafe_o = np.zeros(3 * B, np.float32)
    for t in range(5):                                            # is_safe (hier_decision.py:93-97)
        a = syn.make_actions(rng, 1, 3 * B)[0]
        res = model.rollout_out(a)
        unsafe += res[4]
        unsafe_o += orc.compute_rewards(prev, orc.action_transformation(a), task)[3]     # teacher forced
        prev = res[0].numpy()
    close(unsafe.cpu().numpy(), unsafe_o)
    assert ((unsafe.cpu().numpy() > 0) == (unsafe_o > 0)).all()


@pytest.mark.parametrize('task,V,mode', [('left', 32, 'training'), ('straight', 9, 'selecting'), ('right', 5, 'training'),
                                         ('left', 0, 'training')])
def test_horizon_fused_equals_step_by_step(dm, task, V, mode):
    """ce2e_rollout_horizon (state resident on chip across H steps) is bit-identical to H launches of
    ce2e_rollout_step: every step's five outputs and the final observations."""
    from env_build_b200 import synthetic as syn
    from env_build_b200.rollout import RolloutGraph
    rng = np.random.default_rng(33)
    B, H = 5003, 7
    model = dm.EnvironmentModel(task, mode=mode, veh_mode_list=tiled(task, V))
    model.ref_path.set_path(1)
    ref = syn.make_ref_indexes(rng, B, out_of_range_frac=0.02)
    obs = syn.make_obs(rng, B, task, V, model.ref_path.path_list, ref if mode == 'training' else 1)
    tape = syn.make_actions(rng, H, B)
    outs = []
    for fused in (False, True):
        g = RolloutGraph(model, B, V, H, use_graph=False, fused=fused)
        g.load(obs, ref, tape)
        g.run()
        torch.cuda.synchronize()
        outs.append((g.out5.cpu().numpy().copy(), g.final_obs.numpy().copy()))
    bits_equal(outs[1][0], outs[0][0])
    bits_equal(outs[1][1], outs[0][1])
    with pytest.raises(ValueError):
        m40 = dm.EnvironmentModel(task, mode=mode, veh_mode_list=tiled(task, 40))
        RolloutGraph(m40, 64, 40, 3, use_graph=False, fused=True).run()


def test_free_running_drift_report(dm):
    """free-running H=25 (state fed back on the device, no teacher forcing), B=4096, V=8:
    report the drift against the fp32 oracle and the number of nearest-waypoint flips.  sin/cos
    differ by <= 1.5 ulp per call, so trajectories separate by a few ulp per step; a flip moves the
    reference point by 10 samples (0.33 m) and shows up as a jump in the tracking columns."""
    from env_build_b200 import synthetic as syn
    rng = np.random.default_rng(20210316)
    task, B, V, H = 'left', 4096, 8, 25
    model = dm.EnvironmentModel(task, mode='training')
    paths = model.ref_path.path_list
    ref = syn.make_ref_indexes(rng, B)
    obs = syn.make_obs(rng, B, task, V, paths, ref)
    tape = syn.make_actions(rng, H, B)
    om = orc.EnvironmentModel(task, mode='training', path_list=paths)
    model.reset(obs, ref)
    om.reset(obs, ref)
    worst_ego, worst_veh, flips = 0.0, 0.0, np.zeros(B, bool)
    for t in range(H):
        g = model.rollout_out(tape[t])[0].numpy()
        w = om.rollout_out(tape[t])[0]
        flips |= np.abs(g[:, 6] - w[:, 6]) > 0.05
        keep = ~flips
        worst_ego = max(worst_ego, float(np.abs(g[keep, :6] - w[keep, :6]).max()))
        worst_veh = max(worst_veh, float(np.abs(g[:, 9:] - w[:, 9:]).max()))
    print('free-running H=25, B=%d: max |ego drift| %.3g, max |vehicle drift| %.3g, waypoint flips %d rows (%.3f %%)'
          % (B, worst_ego, worst_veh, int(flips.sum()), 100.0 * flips.mean()))
    assert worst_ego < 2e-4 and worst_veh < 2e-4 and flips.mean() < 0.01


def test_fp64_shadow_accuracy(dm):
    """How far do the fp32 kernel and the fp32 oracle each sit from a float64 evaluation of the same
    op graph (oracle/torch_model.py)?  The kernel must not be further from the exact value than a few
    times the oracle's own rounding error."""
    from env_build_b200 import synthetic as syn
    from oracle import torch_model as tm
    rng = np.random.default_rng(20210317)
    task, B, V = 'left', 3000, 8
    model = dm.EnvironmentModel(task, mode='training')
    paths = model.ref_path.path_list
    ref = syn.make_ref_indexes(rng, B)
    obs = syn.make_obs(rng, B, task, V, paths, ref)
    act = syn.make_actions(rng, 1, B)[0]
    model.reset(obs, ref)
    got = [r.numpy() for r in model.rollout_out(act)]
    om = orc.EnvironmentModel(task, mode='training', path_list=paths)
    om.reset(obs, ref)
    _, margin = om.compute_next_obses(obs, orc.action_transformation(act), return_margin=True)
    want = om.rollout_out(act)
    shadow = tm.rollout_out(torch.tensor(obs, dtype=torch.float64), torch.tensor(act, dtype=torch.float64), task, ref,
                            paths, orc.VEHICLE_MODE_LIST[task])
    ok = margin > 1e-3
    names = ['next_obs', 'rewards', 'punish_train', 'punish_real', 'veh2veh4real', 'veh2road4real']
    for name, g, w, s in zip(names, got, want, shadow):
        s = s.numpy()
        eg = np.abs(g[ok] - s[ok]) / (1.0 + np.abs(s[ok]))
        ew = np.abs(w[ok] - s[ok]) / (1.0 + np.abs(s[ok]))
        print('%-14s max scaled error vs float64: kernel %.3g   oracle %.3g' % (name, eg.max(), ew.max()))
        assert eg.max() <= 4 * ew.max() + 2e-6


def test_closed_loop_policy_in_graph(dm):
    """hier_decision.py:93-96 pattern `actions = policy(obses); model.rollout_out(actions)` captured as one
    CUDA graph: bit-identical to stepping the public API by hand, and within tolerance of the oracle
    driven by the recorded actions."""
    from env_build_b200 import synthetic as syn
    from env_build_b200.rollout import RolloutGraph
    rng = np.random.default_rng(5)
    task, B, V, H = 'left', 1500, 8, 5
    model = dm.EnvironmentModel(task, mode='selecting')
    obs = syn.make_obs(rng, B, task, V, model.ref_path.path_list, 1)
    gen = torch.Generator(device='cuda').manual_seed(3)
    w1 = torch.randn((41, 32), device='cuda', generator=gen) * 0.05
    w2 = torch.randn((32, 2), device='cuda', generator=gen) * 0.3

    def policy(o):
        return torch.tanh(torch.tanh(o @ w1) @ w2)

    model.ref_path.set_path(1)
    g = RolloutGraph(model, B, V, H, policy=policy)
    g.load(obs)
    g.run()
    g.load(obs)
    g.run()                                            # a replay, not the capture pass
    torch.cuda.synchronize()
    tape, out5, final = g.tape.cpu().numpy(), g.out5.cpu().numpy(), g.final_obs.numpy()
    assert np.abs(tape).max() <= 1.0 and np.abs(tape).std() > 0.05
    # by hand through the public API
    model.add_traj(obs, 1)
    for t in range(H):
        a = policy(torch.as_tensor(model.obses))
        bits_equal(a.cpu().numpy(), tape[t])
        res = model.rollout_out(a)
        bits_equal(np.stack([r.numpy() for r in res[1:]]), out5[t])
    bits_equal(res[0].numpy(), final)
    # oracle driven by the recorded actions, re-seeded from the device observations every step
    om = orc.EnvironmentModel(task, mode='selecting', path_list=model.ref_path.path_list)
    model.add_traj(obs, 1)
    cur = obs
    for t in range(H):
        om.add_traj(cur, 1)
        want = om.rollout_out(tape[t])
        res = model.rollout_out(tape[t])
        for a, b in zip(res[1:], want[1:]):
            close(a.numpy(), b)
        close(res[0].numpy()[:, :6], want[0][:, :6])
        cur = res[0].numpy()
    with pytest.raises(ValueError):
        RolloutGraph(model, B, V, H, fused=True, policy=policy)


def test_reentrant_from_two_threads_and_streams(dm):
    """include/ce2e.h: entry points are re-entrant, work goes to the caller's stream.  Two host threads,
    each with its own stream, model and task, step concurrently; each must reproduce its own
    single-threaded results bit for bit."""
    import threading
    from env_build_b200 import synthetic as syn
    jobs = []
    for k, (task, V) in enumerate((('left', 32), ('right', 9))):
        rng = np.random.default_rng(900 + k)
        B = 20000 + 17 * k
        model = dm.EnvironmentModel(task, mode='training', veh_mode_list=tiled(task, V))
        ref = syn.make_ref_indexes(rng, B)
        obs = syn.make_obs(rng, B, task, V, model.ref_path.path_list, ref)
        tape = syn.make_actions(rng, 12, B)
        model.ref_path.handle                                   # tables built up front
        jobs.append(dict(model=model, obs=obs, ref=ref, tape=tape))

    def rollout(job, stream):
        with torch.cuda.stream(stream):
            m = job['model']
            m.reset(job['obs'], job['ref'])
            acc = 0
            for t in range(len(job['tape'])):
                res = m.rollout_out(job['tape'][t])
                acc = acc + torch.stack([r.clone() for r in res[1:]])
            out = (res[0].clone(), acc)
        stream.synchronize()
        return out

    want = [rollout(j, torch.cuda.Stream()) for j in jobs]
    got = [None, None]
    errs = []

    def worker(k):
        try:
            torch.cuda.set_device(0)
            for _ in range(3):
                got[k] = rollout(jobs[k], torch.cuda.Stream())
        except Exception as e:                                  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for g, w in zip(got, want):
        assert torch.equal(g[0], w[0]) and torch.equal(g[1], w[1])


def test_graph_capture_without_warmup(tmp_path):
    """include/ce2e.h: every call is CUDA-graph capturable.  A fresh process captures the fused step, the
    done kernel and the backward kernel in one graph WITHOUT calling them first (only the synchronous
    ce2e_paths_create ran), replays it and compares with the eager results."""
    import subprocess
    import sys
    script = tmp_path / 'capture.py'
    script.write_text('''
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, %r)
from env_build_b200 import _lib, synthetic as syn
from env_build_b200.dynamics_and_models import EnvironmentModel, padded_rows
rng = np.random.default_rng(1)
task, B, V = 'left', 5000, 8
m = EnvironmentModel(task, mode='training')
ref = syn.make_ref_indexes(rng, B)
obs_h = syn.make_obs(rng, B, task, V, m.ref_path.path_list, ref)
dev = torch.device('cuda')
obs = padded_rows(B, 41, 9, dev); obs.copy_(torch.as_tensor(obs_h))
act = torch.as_tensor(syn.make_actions(rng, 1, B)[0], device=dev)
dref = torch.as_tensor(ref, device=dev, dtype=torch.int32)
lib, vp = _lib.load(), (lambda t: ctypes.c_void_p(t.data_ptr()))
h = m.ref_path.handle                                  # synchronous setup, outside the capture

def run(stream):
    nxt = padded_rows(B, 41, 9, dev); out5 = torch.empty((5, B), device=dev); sc = torch.empty((B, 2), device=dev)
    done = torch.empty((B,), dtype=torch.int8, device=dev)
    g_obs = torch.empty((B, 9), device=dev); g_act = torch.empty((B, 2), device=dev)
    ones9 = torch.ones((B, 9), device=dev); ones5 = torch.ones((5, B), device=dev)
    def enqueue():
        s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.ce2e_rollout_step(h, 0, vp(dref), vp(obs), obs.stride(0), vp(act), ctypes.byref(m._turn), V, V, 0,
                                         vp(nxt), nxt.stride(0), vp(out5), vp(sc), B, s))
        _lib.check(lib.ce2e_judge_done(0, vp(nxt), nxt.stride(0), vp(sc), V, 0, 0, vp(done), B, s))
        _lib.check(lib.ce2e_rollout_step_backward(h, 0, vp(dref), vp(obs), obs.stride(0), vp(act), V, 0, vp(ones9), 9,
                                                  vp(ones5), vp(g_obs), 9, vp(g_act), B, s))
    return enqueue, (nxt, out5, done, g_obs, g_act)

enq, outs = run(None)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    enq()
g.replay()
torch.cuda.synchronize()
captured = [o.clone() for o in outs]
enq2, outs2 = run(None)
enq2()
torch.cuda.synchronize()
for a, b in zip(captured, outs2):
    assert torch.equal(a, b)
assert int((captured[2] != 0).sum()) > 0
print('CAPTURE OK')
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, str(script)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    assert r.returncode == 0 and b'CAPTURE OK' in r.stdout, r.stdout.decode()[-3000:]


def test_maximum_vehicle_count(dm):
    """V = CE2E_MAX_VEH = 256 vehicles per row (D = 1033) against the oracle; one more is refused."""
    from env_build_b200 import _lib, synthetic as syn
    rng = np.random.default_rng(256)
    task, B, V = 'straight', 301, _lib.MAX_VEH
    modes = tiled(task, V)
    model = dm.EnvironmentModel(task, mode='training', veh_mode_list=modes)
    ref = syn.make_ref_indexes(rng, B)
    obs = syn.make_obs(rng, B, task, V, model.ref_path.path_list, ref)
    act = syn.make_actions(rng, 1, B)[0]
    model.reset(obs, ref)
    res = model.rollout_out(act)
    om = orc.EnvironmentModel(task, mode='training', path_list=model.ref_path.path_list, veh_mode_list=modes)
    om.reset(obs, ref)
    sc = orc.action_transformation(act)
    want5 = orc.compute_rewards(obs, sc, task)[:5]
    want, margin = om.compute_next_obses(obs, sc, return_margin=True)
    got = res[0].numpy()
    assert got.shape == (B, 9 + 4 * V)
    for a, b in zip(res[1:], want5):
        close(a.numpy(), b)
    close(got[:, :6], want[:, :6])
    close(got[:, 9:], want[:, 9:])
    ok = margin > 1e-4
    close(got[ok, 6:9], want[ok, 6:9])
    with pytest.raises(ValueError):
        big = dm.EnvironmentModel(task, mode='training', veh_mode_list=tiled(task, V + 1))
        big.reset(np.zeros((4, 9 + 4 * (V + 1)), np.float32), np.zeros(4, np.int32))
        big.rollout_out(np.zeros((4, 2), np.float32))


# ------------------------------------------------------------------------------------------
# the warp-pair TMA kernel against the cp.async kernel (same arithmetic, same summation order)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('task,V,n,mode', [('left', 32, 0, 'training'), ('left', 8, 0, 'selecting'),
                                           ('straight', 9, 0, 'training'), ('right', 5, 0, 'training'),
                                           ('left', 1, 0, 'training'), ('left', 2, 3, 'training'),
                                           ('straight', 33, 10, 'selecting'), ('right', 40, 0, 'training'),
                                           ('left', 256, 0, 'training')])
def test_tma_kernel_equals_cp_async_kernel(dm, task, V, n, mode):
    """ce2e_set_tma(1) (k_model_step_pair: 2-D tensor-map loads / stores, warp pair per 32 rows) and
    ce2e_set_tma(0) (k_model_step: cp.async, two lanes per row) must agree bit for bit on every output,
    for ragged batch sizes around the 32-row tile and for odd / tiny / maximal vehicle counts."""
    from env_build_b200 import _lib, synthetic as syn
    rng = np.random.default_rng(1000 + V + n)
    for B in (32, 33, 95, 4097):
        model = dm.EnvironmentModel(task, n, mode=mode, veh_mode_list=tiled(task, V))
        ref = syn.make_ref_indexes(rng, B, out_of_range_frac=0.05)
        obs = syn.make_obs(rng, B, task, V, model.ref_path.path_list, ref if mode == 'training' else 1, n)
        act = syn.make_actions(rng, 2, B)
        res = {}
        for tma in (2, 3, 0):                               # overlapped split, balanced split, cp.async kernel
            old = _lib.set_tma(tma)
            try:
                n0 = _lib.launch_count()
                if mode == 'training':
                    model.reset(obs, ref)
                else:
                    model.add_traj(obs, 1)
                outs = []
                for t in range(2):                          # two steps: the second reads what the first stored
                    outs.append([r.numpy().copy() for r in model.rollout_out(act[t])])
                res[tma] = outs
                assert _lib.launch_count() - n0 == 2
            finally:
                _lib.set_tma(old)
        for tma in (2, 3):
            for a, b in zip(res[tma], res[0]):
                for x, y in zip(a, b):
                    bits_equal(x, y)


def test_tma_kernel_is_the_default_path(dm):
    """The fused step on padded rows must launch k_model_step_pair (checked through the kernel's name in
    a profiler-free way: the two kernels need different dynamic shared memory, which the library reports
    through ce2e_last_step_kernel)."""
    from env_build_b200 import _lib, synthetic as syn
    rng = np.random.default_rng(5)
    model = dm.EnvironmentModel('left', mode='training', veh_mode_list=tiled('left', 32))
    ref = syn.make_ref_indexes(rng, 256)
    obs = syn.make_obs(rng, 256, 'left', 32, model.ref_path.path_list, ref)
    model.reset(obs, ref)
    model.rollout_out(syn.make_actions(rng, 1, 256)[0])
    assert _lib.load().ce2e_last_step_kernel() == 2           # 2 = k_model_step_pair (TMA)
    old = _lib.set_tma(False)
    try:
        model.rollout_out(syn.make_actions(rng, 1, 256)[0])
        assert _lib.load().ce2e_last_step_kernel() == 1       # 1 = k_model_step (cp.async)
    finally:
        _lib.set_tma(old)


def test_config_set_changes_only_its_term(dm):
    """CrossroadConfig / ce2e_config_set (the reference's baked-in constants as one frozen
    record).  Doubling one reward weight (DM:297-298) doubles that scaled term of the reward dict and
    changes nothing else; a narrower lane (EU:15) moves the road terms and leaves the vehicle terms alone;
    set_config(None) restores the reference's values bit for bit."""
    from env_build_b200 import synthetic as syn
    from env_build_b200.endtoend_env_utils import CrossroadConfig, set_config, get_config
    rng = np.random.default_rng(9)
    task, B, V = 'straight', 2000, 9
    model = dm.EnvironmentModel(task, mode='selecting')
    obs = syn.make_obs(rng, B, task, V, model.ref_path.path_list, 1)
    sc = orc.action_transformation(syn.make_actions(rng, 1, B)[0])

    def rewards():
        r = model.compute_rewards(obs, sc)
        return [x.numpy() for x in r[:5]], {k: v.numpy() for k, v in r[5].items()}
    base5, base_d = rewards()
    assert get_config() == CrossroadConfig()
    try:
        set_config(CrossroadConfig(w_devi_y=1.6))
        got5, got_d = rewards()
        bits_equal(got_d['scaled_devi_y'], np.float32(2) * base_d['scaled_devi_y'])
        for k in base_d:
            if k != 'scaled_devi_y':
                bits_equal(got_d[k], base_d[k])
        for i in range(1, 5):
            bits_equal(got5[i], base5[i])
        assert not np.array_equal(got5[0], base5[0])
        set_config(CrossroadConfig(LANE_WIDTH=3.5))
        got5, got_d = rewards()
        bits_equal(got_d['veh2veh4training'], base_d['veh2veh4training'])
        bits_equal(got_d['veh2veh4real'], base_d['veh2veh4real'])
        bits_equal(got5[0], base5[0])
        assert not np.array_equal(got_d['veh2road4training'], base_d['veh2road4training'])
    finally:
        set_config(None)
    got5, got_d = rewards()
    for a, b in zip(got5, base5):
        bits_equal(a, b)
    # the fused step sees the same constant block
    model.add_traj(obs, 1)
    r0 = model.rollout_out(syn.make_actions(np.random.default_rng(1), 1, B)[0])[1].numpy()
    try:
        set_config(CrossroadConfig(w_punish_steer=10.))
        model.add_traj(obs, 1)
        r1 = model.rollout_out(syn.make_actions(np.random.default_rng(1), 1, B)[0])[1].numpy()
    finally:
        set_config(None)
    assert not np.array_equal(r0, r1)
