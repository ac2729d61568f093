"""The oracle (oracle/crossroad_oracle.py) against the golden vectors produced by the
UNMODIFIED reference source (tests/golden/make_golden.py).  Bit-equality is required:
both sides do fp32 IEEE element-wise arithmetic with float64-rounded transcendentals,
so any difference is a transcription error in the oracle's expression trees."""
import numpy as np
import pytest

from conftest import TASKS, golden_paths
from oracle import crossroad_oracle as orc


def _same(a, b, what=''):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.dtype.kind == 'f':
        assert a.dtype == b.dtype == np.float32, (what, a.dtype, b.dtype)
        ok = (a.view(np.int32) == b.view(np.int32)) | (np.isnan(a) & np.isnan(b)) | ((a == 0) & (b == 0))
        assert ok.all(), (what, int((~ok).sum()), a[~ok][:5], b[~ok][:5])
    else:
        assert (a == b).all(), what


def test_constants(golden_common):
    c = golden_common
    assert (orc.L, orc.W, orc.LANE_WIDTH, orc.LANE_NUMBER, orc.CROSSROAD_SIZE, orc.EXPECTED_V) == \
        tuple(float(c['const_' + k]) for k in ('L', 'W', 'LANE_WIDTH', 'LANE_NUMBER', 'CROSSROAD_SIZE', 'EXPECTED_V'))
    for task in TASKS:
        assert list(c['mode_list_' + task]) == orc.VEHICLE_MODE_LIST[task]
    for k, v in orc.VEHICLE_PARAMS.items():
        assert float(c['vp_' + k]) == v


def test_f_xu(golden_common):
    c = golden_common
    nxt, par = orc.f_xu(c['fxu_states'], c['fxu_actions'], 0.1)
    _same(nxt, c['fxu_next'], 'f_xu next')
    _same(par, c['fxu_params'], 'f_xu params')
    nxt, par = orc.prediction(c['fxu_states'], c['fxu_actions'], 10)
    _same(nxt, c['pred_next'], 'prediction next')
    _same(par, c['pred_params'], 'prediction params')


def test_f_xu_hand_kat():
    # straight-line motion at heading 90 deg
    nxt, par = orc.f_xu(np.array([[5, 0, 0, 0, 0, 90]], np.float32), np.zeros((1, 2), np.float32), 0.1)
    assert nxt[0, 0] == 5 and nxt[0, 1] == 0 and nxt[0, 2] == 0 and nxt[0, 4] == 0.5 and nxt[0, 5] == 90
    assert abs(nxt[0, 3] - 0.5 * np.cos(np.float64(np.float32(np.pi) / np.float32(2)))) < 1e-12 + 1e-7 * 2.2e-8


def test_action_scaling(golden_common):
    c = golden_common
    _same(orc.action_transformation(c['act_norm']), c['act_scaled'])
    s = orc.action_transformation(np.array([[1, 1], [-1, -1], [2, 0]], np.float32))
    assert np.allclose(s, [[0.4, 1.5], [-0.4, -3.0], [0.42, -0.75]], atol=1e-6)


def test_phi_wrap(golden_common):
    _same(orc.deal_with_phi_diff(golden_common['phidiff_in']), golden_common['phidiff_out'])


def test_predict_for_a_mode(golden_common):
    c = golden_common
    for mode in ('dl', 'rd', 'ur', 'lu', 'dr', 'ru', 'ul', 'ld', 'du', 'ud', 'lr', 'rl'):
        _same(orc.predict_for_a_mode(c['pfm_in'], mode), c['pfm_out_' + mode], mode)


def test_gym_helpers(golden_common):
    c = golden_common
    for task in TASKS:
        got = np.array([orc.judge_feasible(x, y, task) for x, y in c['jf_xy']])
        assert (got == c['jf_' + task]).all()
    assert [orc.deal_with_phi(float(p)) for p in c['dwp_in']] == list(c['dwp_out'])


@pytest.mark.parametrize('task', TASKS)
def test_path_tables(task, golden_task):
    g = golden_task(task)
    paths, len_list, ctrl = orc.construct_ref_paths(task)
    assert (np.array(len_list) == g['path_len_list']).all()
    assert np.array_equal(np.array(ctrl, dtype=np.float64), g['control_points'])
    expect_len = dict(left=3657, straight=3897, right=3117)[task]
    for i, p in enumerate(paths):
        assert len(p[0]) == expect_len
        _same(p[0], g['path%d_x' % i], 'x')
        _same(p[1], g['path%d_y' % i], 'y')
        # heading: the reference calls NumPy's fp32 arctan2 (build/SIMD dependent last ulp);
        # the oracle uses float64 atan2 rounded to fp32 (machine independent).  NumPy SVML atan2f
        # is a <= 4 ulp routine; measured here: max 3 ulp (relative 3.6e-7, far inside rtol 1e-5).
        a, b = p[2], g['path%d_phi' % i]
        ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
        ulp[(a == 0) & (b == 0)] = 0
        assert ulp.max() <= 4, (task, i, int(ulp.max()))


@pytest.mark.parametrize('task', TASKS)
def test_tracking(task, golden_task):
    g = golden_task(task)
    rp = orc.ReferencePath(task, 0, path_list=golden_paths(g))      # the reference's own tables
    for pi in range(3):
        rp.set_path(pi)
        xs, ys, phis, vs = g['trk%d_in' % pi].T
        for n in (0, 3, 10):
            _same(rp.tracking_error_vector(xs, ys, phis, vs, n), g['trk%d_n%d' % (pi, n)], 'trk n=%d' % n)
        idx, pts = rp.find_closest_point(xs, ys)
        assert (idx == g['fcp%d_idx' % pi]).all()
        _same(np.stack(pts, 1), g['fcp%d_pts' % pi])
        idx5, _ = rp.find_closest_point(xs, ys, ratio=5)
        assert (idx5 == g['fcp%d_idx_r5' % pi]).all()
        fut = rp.future_n_data(np.array([600, 0, 3500, len(rp.path[0]) - 3], np.int64), 5)
        _same(np.stack([np.stack(f, 1) for f in fut], 0), g['fut%d' % pi])
        # on-waypoint poses -> zero tracking error (SURVEY 8c KAT)
        t0 = rp.tracking_error_vector(xs, ys, phis, vs, 0)
        assert np.abs(t0[8:12]).max() == 0


def test_tracking_reference_script_inputs(golden_task):
    """The inputs of the reference's test_tracking_error_vector (DM:805-808), task left path 0;
    expected values from SURVEY 8c (scratch restatement) to 5 decimals."""
    g = golden_task('left')
    want = np.array([[-0.32387, -7.79610, 2], [-8.86666, 2.28967, 4], [6.62911, -0.16386, 2], [2.51011, 8.17429, 2]])
    assert np.allclose(g['trk0_n0'][:4], want, atol=1e-5)


@pytest.mark.parametrize('task', TASKS)
def test_compute_rewards(task, golden_task):
    g = golden_task(task)
    keys = list(g['rew_dict_keys'])
    for V in (orc.VEH_NUM[task], 32):
        ob, an = g['rew_V%d_obs' % V], g['rew_V%d_act' % V]
        r = orc.compute_rewards(ob, orc.action_transformation(an), task, 0)
        _same(np.stack(r[:5], 1), g['rew_V%d_out5' % V], 'out5 V=%d' % V)
        _same(np.stack([r[5][k] for k in keys], 1), g['rew_V%d_dict' % V], 'dict V=%d' % V)
        assert sorted(orc.REWARD_DICT_KEYS) == keys
        # pad vehicles far away -> veh2veh exactly zero
        assert r[3][0] == 0
        assert (g['rew_V%d_out5' % V][:, 1] > 0).any(), 'synthetic batch should exercise the hinge terms'


@pytest.mark.parametrize('task', TASKS)
@pytest.mark.parametrize('tag', ['cfg1', 'selV', 'trnV', 'sel32', 'trn32', 'seln10', 'trnn3'])
def test_rollout(task, tag, golden_task):
    """H=25 free-running rollout_out, bit-exact against the reference (config #1 = 'cfg1')."""
    g = golden_task(task)
    ob0, ref, tape = g['ro_%s_obs0' % tag], g['ro_%s_ref' % tag], g['ro_%s_tape' % tag]
    n = {'seln10': 10, 'trnn3': 3}.get(tag, 0)
    V = (ob0.shape[1] - 6 - 3 * (n + 1)) // 4
    mode = 'training' if tag.startswith('trn') else 'selecting'
    from env_build_b200.synthetic import tiled_mode_list
    model = orc.EnvironmentModel(task, n, mode=mode, path_list=golden_paths(g),
                                 veh_mode_list=tiled_mode_list(orc.VEHICLE_MODE_LIST[task], V))
    if mode == 'training':
        model.reset(ob0, ref)
    else:
        model.add_traj(ob0, int(g['ro_%s_path' % tag]))
    for t in range(tape.shape[0]):
        res = model.rollout_out(tape[t])
        _same(res[0], g['ro_%s_obs' % tag][t], '%s obs t=%d' % (tag, t))
        _same(np.stack(res[1:], 1), g['ro_%s_out5' % tag][t], '%s out5 t=%d' % (tag, t))
    if tag in ('selV', 'sel32'):
        model.add_traj(ob0, int(g['ro_%s_path' % tag]))
        _same(model.ss(ob0, tape[0], lam=0.1), g['ss_%s' % tag], 'ss')
    if mode == 'training':
        # ref_index 3 matches no path -> zero tracking columns (DM:342-353)
        assert ref[0] == 3 and np.abs(g['ro_%s_obs' % tag][:, 0, 6:9 + 3 * n]).max() == 0


@pytest.mark.parametrize('task', TASKS)
def test_gym_side_against_reference_methods(task):
    """gym_next_ego_state / gym_ego_dynamics / gym_judge_done against the UNMODIFIED reference
    methods (CrossroadEnd2end._get_next_ego_state, _get_ego_dynamics, _judge_done, compute_reward,
    Traffic.collision_check) run by tests/golden/make_golden_env.py."""
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, 'env_%s.npz' % task), allow_pickle=False))
    scaled, nxt, params = orc.gym_next_ego_state(g['obs'][:, :6], g['act'])
    _same(scaled, g['scaled'], 'scaled action')
    _same(nxt, g['next_ego'], 'next ego state')
    _same(params, g['params'], 'tyre params')
    corners, r_bound = orc.gym_ego_dynamics(nxt, params)
    # the reference mixes np.float32 scalars with Python floats here, so its own result depends on
    # the NumPy version (float64 before NEP 50, float32 after); the goldens were made with NumPy 2
    assert np.allclose(corners, g['corners'], rtol=0, atol=1e-5)
    fin = np.isfinite(g['r_bound'])
    assert np.allclose(r_bound[fin], g['r_bound'][fin], rtol=1e-6) and (np.isfinite(r_bound) == fin).all()
    rew = orc.compute_rewards(g['obs'], scaled, task)[0]
    _same(rew, g['reward'], 'reward')
    code, margin = orc.gym_judge_done(nxt, params, g['obs'][:, 6], g['veh_after'], task, g['v_light'])
    assert (code == g['done_code']).all(), np.flatnonzero(code != g['done_code'])
    assert set(np.unique(code).tolist()) >= {0, 1, 2, 3, 4, 6}
    if task != 'right':
        assert (code == 5).any()


@pytest.mark.parametrize('task', TASKS)
def test_vehicle_selection_against_reference_method(task):
    """select_interested_vehicles against the UNMODIFIED CrossroadEnd2end._construct_veh_vector_short
    (E2E:340-464) on 300 random scenes per task (half-metre coordinates: many equal sort keys,
    empty scenes, red-light virtual vehicles)."""
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, 'env_%s.npz' % task), allow_pickle=False))
    for i in range(len(g['sel_out'])):
        got = orc.select_interested_vehicles(g['sel_veh'][i], g['sel_cls'][i], g['sel_ego'][i, 0], g['sel_ego'][i, 1],
                                             task, int(g['sel_light'][i]), bool(g['sel_virtual'][i]))
        _same(got, g['sel_out'][i], 'scene %d' % i)


@pytest.mark.parametrize('task', TASKS)
def test_reset_init_state_against_reference_method(task, golden_task):
    """oracle.reset_init_state == the UNMODIFIED CrossroadEnd2end._reset_init_state (E2E:472-499) fed the
    same two np.random.random() draws (make_golden_env.py replays them by seed)."""
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, 'env_%s.npz' % task), allow_pickle=False))
    paths = golden_paths(golden_task(task))
    for u, ego, k in zip(g['reset_u'], g['reset_ego'], g['reset_path']):
        _same(orc.reset_init_state(task, paths[int(k)], float(u[0]), float(u[1])), ego, 'reset ego')
    idx = (g['reset_u'][:, 0] * orc.RESET_SPAN[task]).astype(int) + 700
    assert idx.min() >= 700 and idx.max() < 700 + orc.RESET_SPAN[task]


def test_philox_known_answers():
    """Random123's known-answer vectors for Philox4x32-10 (kat_vectors: philox4x32 10)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = orc.philox4x32_10(np.array([c], np.uint32), np.array([k], np.uint32))[0]
        assert tuple(int(x) for x in got) == want


def test_env_reset_rows_distribution():
    """The restated device reset: ego on a waypoint inside the reference's start window with v in [0, 8),
    zero tracking error up to the path's own granularity, collision-free start, deterministic in
    (seed, row, episode) and independent of which other rows are drawn with it."""
    task = 'left'
    paths = orc.construct_ref_paths(task)[0]
    rows = np.arange(2000)
    obs, ref, red = orc.env_reset_rows(1234, rows, np.zeros(2000, int), task, paths, 8)
    assert obs.dtype == np.float32 and obs.shape == (2000, 41) and set(np.unique(ref)) == {0, 1, 2}
    assert (obs[:, 0] >= 0).all() and (obs[:, 0] < 8).all() and (obs[:, 1:3] == 0).all()
    assert abs(obs[:, 0].mean() - 4.0) < 0.2 and 0.05 < red.mean() < 0.15
    assert np.abs(obs[:, 6]).max() < 0.1          # the projection lands on every 10th waypoint (DM:704-714)
    assert (obs[:, 8] == obs[:, 0] - np.float32(8)).all()
    veh = obs[:, 9:].reshape(2000, 8, 4)
    d = np.hypot(veh[:, :, 0] - obs[:, None, 3], veh[:, :, 1] - obs[:, None, 4])
    assert d.min() >= 6.0 - 1e-3 and (veh[:, :, 2] >= 0).all() and (veh[:, :, 2] < 8).all()
    assert (veh[:, :, 3] > -180).all() and (veh[:, :, 3] <= 180).all()
    sub = np.array([5, 77, 1999])
    o2, r2, _ = orc.env_reset_rows(1234, sub, np.zeros(3, int), task, paths, 8)
    _same(o2, obs[sub])
    assert (r2 == ref[sub]).all()
    o3, _, _ = orc.env_reset_rows(1234, sub, np.ones(3, int), task, paths, 8)
    assert not np.array_equal(o3, o2)


@pytest.mark.parametrize('task', TASKS)
def test_gradient_golden_matches_torch_oracle(task):
    """tests/golden/grad_<task>.npz (vector-Jacobian product of the UNMODIFIED reference rollout_out under
    the shim's GradientTape, TensorFlow's autodiff rules) against torch.autograd on the float64 restatement
    in oracle/torch_model.py: two independent derivations of the same gradient."""
    import os
    import torch
    from conftest import GOLDEN
    from oracle import torch_model as tm
    g = dict(np.load(os.path.join(GOLDEN, 'grad_%s.npz' % task), allow_pickle=False))
    paths = orc.construct_ref_paths(task)[0]
    obs = torch.tensor(g['obs'], dtype=torch.float64, requires_grad=True)
    act = torch.tensor(g['act'], dtype=torch.float64, requires_grad=True)
    res = tm.rollout_out(obs, act, task, g['ref'], paths, orc.VEHICLE_MODE_LIST[task])
    loss = (res[0][:, :9] * torch.tensor(g['g_next9'], dtype=torch.float64)).sum() + \
        (torch.stack(res[1:]) * torch.tensor(g['g_out5'], dtype=torch.float64)).sum()
    loss.backward()
    om = orc.EnvironmentModel(task, mode='training', path_list=paths)
    om.reset(g['obs'], g['ref'])
    _, margin = om.compute_next_obses(g['obs'], orc.action_transformation(g['act']), return_margin=True)
    ok = margin > 1e-3                    # same closest waypoint in fp32 and float64
    assert ok.mean() > 0.95
    # the golden's forward ran in fp32, the oracle's in float64: gradients agree to fp32 forward accuracy
    assert np.allclose(obs.grad.numpy()[ok][:, :9], g['grad_obs9'][ok], rtol=2e-4, atol=2e-4)
    assert np.allclose(act.grad.numpy()[ok], g['grad_act'][ok], rtol=2e-4, atol=2e-4)
    assert np.abs(obs.grad.numpy()[:, 9:]).max() == 0
