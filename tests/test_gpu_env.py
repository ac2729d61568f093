"""Batched SUMO-free CrossroadEnd2end (env_build_b200/endtoend.py -> ce2e_env_step) against the
oracle's restatement of endtoend.py:132-256 / traffic.py:263-295.  Needs a GPU.
Done codes are integers: they must match exactly on every row whose deciding comparison has a
slack above 1e-3 (the oracle evaluates the predicates in float64 like the reference's Python
scalars, the kernel in fp32)."""
import numpy as np
import pytest
import torch

from conftest import TASKS
from oracle import crossroad_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def e2e():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from env_build_b200 import _lib
    if _lib.needs_build():
        _lib.build()
    from env_build_b200 import endtoend
    return endtoend


def craft(obs, task, paths):
    """Rows that exercise the done types (collisions come from the synthetic near vehicles)."""
    o = obs.copy()
    q = len(o) // 8
    o[q:6 * q, 9::4] = 400.0                                    # no vehicles around these rows
    o[2 * q:3 * q, 2] = 3.0                                     # large yaw rate -> break_stability
    o[2 * q:3 * q, 0] = 9.0
    goal = dict(left=(-36.0, 5.6, 180.0), straight=(5.6, 36.0, 90.0), right=(36.0, -5.6, 0.0))[task]
    o[3 * q:4 * q, 3], o[3 * q:4 * q, 4], o[3 * q:4 * q, 5] = goal   # beyond the goal line -> good_done
    o[3 * q:4 * q, 1:3] = 0
    off = dict(left=(20.0, -40.0, 90.0), straight=(-10.0, -40.0, 90.0), right=(-3.0, -40.0, 90.0))[task]
    o[4 * q:5 * q, 3], o[4 * q:5 * q, 4], o[4 * q:5 * q, 5] = off    # off the approach lanes -> road constraint
    o[5 * q:6 * q, 3] += 25.0                                   # far from the path -> road / deviate
    return o


@pytest.mark.parametrize('task', TASKS)
def test_env_step_matches_oracle(e2e, task):
    from env_build_b200 import synthetic as syn
    rng = np.random.default_rng(20210314)
    B, V = 4000, orc.VEH_NUM[task]
    env = e2e.CrossroadEnd2end(task, num_envs=B)
    env.seed(3)
    env.reset()
    paths = env.ref_path.path_list
    ref = syn.make_ref_indexes(rng, B)
    obs = craft(syn.make_obs(rng, B, task, V, paths, ref), task, paths)
    act = syn.make_actions(rng, 1, B)[0]
    env.obs = env.env_model._adopt(obs)
    env.ref_indexes = torch.as_tensor(ref, device='cuda')
    got_obs, got_rew, got_done, info = env.step(act)
    want_obs, want_rew, want_code, margin = orc.gym_env_step(obs, act, task, ref, paths, orc.VEHICLE_MODE_LIST[task])
    g = got_obs.numpy()
    assert np.allclose(g[:, :6], want_obs[:, :6], rtol=1e-5, atol=1e-5)
    assert np.allclose(g[:, 9:], want_obs[:, 9:], rtol=1e-5, atol=1e-5)
    assert np.allclose(got_rew.numpy(), want_rew, rtol=1e-5, atol=1e-5)
    # heading wrapped to (-180, 180], v_x floored at 0 (E2E:281-282)
    assert g[:, 5].max() <= 180.0 and g[:, 5].min() > -180.0 and g[:, 0].min() >= 0.0
    code = info['done_code'].numpy()
    ok = margin > 1e-3
    assert ok.mean() > 0.97
    assert (code[ok] == want_code[ok]).all(), np.flatnonzero(code[ok] != want_code[ok])[:10]
    assert (got_done.numpy() == (code != 0)).all()
    seen = set(np.unique(want_code[ok]).tolist())
    assert {0, 1, 2, 4, 6} <= seen, seen                       # every branch but red light / deviate exercised here


def test_deviate_and_red_light(e2e):
    """|delta_y| > 15 (E2E:223-225) and the red-light rule (E2E:244-245) need dedicated rows."""
    from env_build_b200 import synthetic as syn
    rng = np.random.default_rng(5)
    task, B, V = 'straight', 512, 9
    env = e2e.CrossroadEnd2end(task, num_envs=B)
    env.reset()
    paths = env.ref_path.path_list
    ref = np.zeros(B, np.int32)
    obs = syn.make_obs(rng, B, task, V, paths, ref, edge_frac=0.0)
    obs[:, 9::4] = 400.0
    obs[:, 3] = paths[0][0][2000] + rng.uniform(-1, 1, B)
    obs[:B // 2, 3] -= 20.0                                          # inside the box, 20 m left of the path
    obs[:, 4] = rng.uniform(-20, 20, B)
    obs[:, 5] = 90.0
    obs[:, 1:3] = 0.0
    act = np.zeros((B, 2), np.float32)
    for v_light in (0, 1):
        env.obs = env.env_model._adopt(obs)
        env.ref_indexes = torch.as_tensor(ref, device='cuda')
        env.v_light = v_light
        _, _, _, info = env.step(act)
        _, _, want, margin = orc.gym_env_step(obs, act, task, ref, paths, orc.VEHICLE_MODE_LIST[task], v_light=v_light)
        ok = margin > 1e-3
        assert (info['done_code'].numpy()[ok] == want[ok]).all()
        assert (want[ok] == 3).any() and ((want[ok] == 5).any() == bool(v_light))


def test_single_env_gym_api(e2e):
    env = e2e.CrossroadEnd2end('left')
    env.seed(0)
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (41,) and obs.dtype == np.float32
    assert env.observation_space.shape == (41,) and env.action_space.shape == (2,)
    total, steps, done = 0.0, 0, 0
    while not done and steps < 200:                              # README: 200-step episode cap
        a = np.array([0.0, 0.3], np.float32)
        prev = obs
        obs, reward, done, info = env.step(a)
        r2, rd = env.compute_reward(prev, env._action_transformation_for_end2end(a))
        assert abs(r2 - reward) <= 1e-5 + 1e-5 * abs(reward)
        assert set(rd) == set(info['reward_info']) - {'final_rew'}
        assert isinstance(reward, float) and done in (0, 1) and info['done_type'] in e2e.DONE_TYPES
        total += reward
        steps += 1
    assert steps >= 1 and np.isfinite(total)
    nxt, par = env._get_next_ego_state(np.array([0.0, 0.0], np.float32))
    assert nxt.shape == (6,) and par.shape == (4,)


def test_auto_reset(e2e):
    env = e2e.CrossroadEnd2end('right', num_envs=2048, auto_reset=True)
    env.seed(1)
    env.reset()
    rng = np.random.default_rng(0)
    n_done = 0
    for _ in range(30):
        obs, rew, done, info = env.step(rng.uniform(-1, 1, (2048, 2)).astype(np.float32))
        n_done += int(done.sum())
        o = obs.numpy()
        assert np.isfinite(o).all()
        # rows that were reset start on their path again: tracking error is that of a fresh pose
        assert np.abs(o[done.numpy(), 6]).max(initial=0.0) < 1.0
    assert n_done > 0


@pytest.mark.parametrize('task', TASKS)
def test_judge_done_against_reference_methods(e2e, task):
    """k_env_done on the inputs of the golden vectors made by the UNMODIFIED reference methods
    (tests/golden/make_golden_env.py): next ego state bit-exact except x, y (sin/cos), done codes
    equal wherever the oracle's decision slack exceeds 1e-3."""
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, 'env_%s.npz' % task), allow_pickle=False))
    N, V = g['obs'].shape[0], orc.VEH_NUM[task]
    env = e2e.CrossroadEnd2end(task, num_envs=N)
    nxt, par = env.dynamics.prediction(g['obs'][:, :6], g['scaled'], 10)
    nxt = nxt.numpy()
    assert np.allclose(nxt[:, 1:5], g['next_ego'][:, 1:5], rtol=1e-5, atol=1e-5)
    after = np.concatenate([g['next_ego'], g['obs'][:, 6:9], g['veh_after'].reshape(N, -1)], 1).astype(np.float32)
    _, margin = orc.gym_judge_done(g['next_ego'], g['params'], g['obs'][:, 6], g['veh_after'], task, g['v_light'])
    codes = {}
    for vl in (0, 1):
        env.v_light = vl
        codes[vl] = env._judge_done(after, g['scaled'])[0].numpy()
    got = np.where(g['v_light'] != 0, codes[1], codes[0])
    ok = margin > 1e-3
    assert ok.mean() > 0.95 and (got[ok] == g['done_code'][ok]).all()


@pytest.mark.parametrize('task', TASKS)
def test_vehicle_selection(e2e, task):
    """k_select_vehicles against the goldens of the UNMODIFIED _construct_veh_vector_short
    (bit-exact: the outputs are copies of inputs or fill constants) and, on 20000 random scenes,
    against the oracle."""
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, 'env_%s.npz' % task), allow_pickle=False))
    env = e2e.CrossroadEnd2end(task, num_envs=len(g['sel_out']))
    got = np.zeros_like(g['sel_out'])
    for vl in (0, 1):                       # v_light is a scalar of the call: run both, pick per scene
        env.v_light = vl
        res = env.construct_veh_vectors(g['sel_veh'], g['sel_cls'], g['sel_ego'], g['sel_virtual']).numpy()
        m = (g['sel_light'] != 0) == bool(vl)
        got[m] = res[m]
    assert np.array_equal(got.view(np.int32), g['sel_out'].view(np.int32))
    # larger random batch against the oracle
    rng = np.random.default_rng(8)
    S, N = 20000, 24
    veh = np.stack([np.round(rng.uniform(-45, 45, (S, N)) * 4) / 4, np.round(rng.uniform(-60, 50, (S, N)) * 4) / 4,
                    rng.uniform(0, 8, (S, N)), rng.choice([0., 90., 180., -90.], (S, N))], 2).astype(np.float32)
    cls = rng.integers(-1, 12, (S, N)).astype(np.int8)
    ego = np.stack([rng.uniform(-30, 12, S), rng.uniform(-60, 30, S)], 1).astype(np.float32)
    virt = rng.random(S) < 0.3
    env.v_light = 0
    res = env.construct_veh_vectors(veh, cls, ego, virt).numpy()
    for i in rng.choice(S, 1500, replace=False):
        want = orc.select_interested_vehicles(veh[i], cls[i], ego[i, 0], ego[i, 1], task, 0, bool(virt[i]))
        assert np.array_equal(res[i].view(np.int32), want.view(np.int32)), i
    # the single-env method with the reference's list-of-dicts format
    one = e2e.CrossroadEnd2end(task)
    one.seed(17)
    one.reset()
    names = {c: r for c, r in zip(e2e.ROUTE_CLASSES, [('1o', '4i'), ('1o', '3i'), ('1o', '2i'), ('2o', '1i'), ('2o', '4i'),
                                                       ('2o', '3i'), ('3o', '2i'), ('3o', '1i'), ('3o', '4i'), ('4o', '3i'),
                                                       ('4o', '2i'), ('4o', '1i')])}
    assert [e2e.route_class(names[c]) for c in e2e.ROUTE_CLASSES] == list(range(12))
    assert e2e.route_class(('2o', '1i'), 'R') == e2e.ROUTE_CLASSES.index('dl') and e2e.route_class(None) == -1
    one.all_vehicles = [dict(x=float(v[0]), y=float(v[1]), v=float(v[2]), phi=float(v[3]), route=names[e2e.ROUTE_CLASSES[c]])
                        for v, c in zip(veh[0], cls[0]) if c >= 0]
    ex, ey = one.obs.numpy()[0, 3:5]
    # (a reset leaves the virtual red-light vehicles on for one episode in ten, E2E:120-124: tell the oracle)
    for virt1 in (one.virtual_red_light_vehicle, not one.virtual_red_light_vehicle):
        one.virtual_red_light_vehicle = virt1
        want = orc.select_interested_vehicles(veh[0][cls[0] >= 0], cls[0][cls[0] >= 0], ex, ey, task, int(one.v_light), virt1)
        assert np.array_equal(one._construct_veh_vector_short().view(np.int32), want.view(np.int32))


@pytest.mark.parametrize('V', [1, 5, 9, 32, 37])
def test_judge_done_tiled_equals_scalar(e2e, V):
    """ce2e_judge_done has two kernels (16 B aligned vehicle block -> tiled, else one thread per row);
    the done codes are integers: identical."""
    import ctypes
    from env_build_b200 import _lib, synthetic as syn
    from env_build_b200.dynamics_and_models import padded_rows, build_path_tables
    rng = np.random.default_rng(200 + V)
    task, B = 'left', 2093                                       # ragged: last tile has 13 rows
    paths = build_path_tables(task)[0]
    obs = craft(syn.make_obs(rng, B, task, V, paths, syn.make_ref_indexes(rng, B)), task, paths)
    D = obs.shape[1]
    sc = torch.as_tensor(rng.uniform(-3, 1.5, (B, 2)).astype(np.float32), device='cuda')
    lib, vp = _lib.load(), (lambda t: ctypes.c_void_p(t.data_ptr()))
    codes = []
    for padded in (True, False):
        if padded:
            o = padded_rows(B, D, 9, torch.device('cuda'))
            o.copy_(torch.as_tensor(obs))
        else:
            o = torch.as_tensor(obs, device='cuda').contiguous()
            assert o.stride(0) % 4 != 0 or (o.data_ptr() + 36) % 16 != 0
        done = torch.full((B,), -1, dtype=torch.int8, device='cuda')
        _lib.check(lib.ce2e_judge_done(0, vp(o), o.stride(0), vp(sc), V, 0, 0, vp(done), B, None))
        torch.cuda.synchronize()
        codes.append(done.cpu().numpy())
    assert (codes[0] >= 0).all() and (codes[0] == codes[1]).all()
    assert (codes[0] == 1).sum() > 0 or V < 5                    # collisions present


# ------------------------------------------------------------------------------------------
# on-device reset (ce2e_env_reset), graph stepping, set_traj
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('task,V', [('left', 8), ('straight', 9), ('right', 5), ('left', 32)])
def test_device_reset_matches_oracle(e2e, task, V):
    """reset() on the device against the NumPy restatement of the same draws (oracle.env_reset_rows, whose
    ego part is pinned to the unmodified _reset_init_state): every column bit for bit."""
    B = 3000
    env = e2e.CrossroadEnd2end(task, num_envs=B, veh_num=V)
    env.seed(20210318)
    obs = env.reset().numpy()
    want, ref, red = orc.env_reset_rows(20210318, np.arange(B), np.zeros(B, int), task, env.ref_path.path_list, V)
    assert (env.ref_indexes.cpu().numpy() == ref).all() and set(np.unique(ref)) == {0, 1, 2}
    assert np.array_equal(obs.view(np.int32), want.view(np.int32)), np.argwhere(obs != want)[:5]
    assert (env._bufs['red'].cpu().numpy().astype(bool) == red).all()
    # a second reset is the next episode of every row
    obs2 = env.reset().numpy()
    want2, _, _ = orc.env_reset_rows(20210318, np.arange(B), np.ones(B, int), task, env.ref_path.path_list, V)
    assert np.array_equal(obs2.view(np.int32), want2.view(np.int32))
    # pinned path
    obs3 = env.reset(ref_index=2).numpy()
    want3, ref3, _ = orc.env_reset_rows(20210318, np.arange(B), np.full(B, 2), task, env.ref_path.path_list, V, fixed_path=2)
    assert (ref3 == 2).all() and np.array_equal(obs3.view(np.int32), want3.view(np.int32))


def test_auto_reset_on_device_in_a_graph(e2e):
    """auto_reset with use_graph=True: step = fused model step + done kernel + reset kernel replayed as one
    CUDA graph; identical to the eager path, rows that finish are replaced by the oracle's reset rows of
    their next episode, and nothing synchronises with the host (no nonzero / item in step)."""
    task, B, V = 'left', 4096, 8
    envs = [e2e.CrossroadEnd2end(task, num_envs=B, veh_num=V, auto_reset=True, use_graph=g, reward_info=False)
            for g in (True, False)]
    for env in envs:
        env.seed(7)
        env.reset()
    rng = np.random.default_rng(0)
    episodes = np.zeros(B, int)
    n_done = 0
    for t in range(40):
        act = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
        outs = [env.step(act) for env in envs]
        o_g, o_e = outs[0][0].numpy(), outs[1][0].numpy()
        assert np.array_equal(o_g.view(np.int32), o_e.view(np.int32)), t
        assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
        done = outs[0][2].cpu().numpy()
        n_done += int(done.sum())
        episodes += done
        rows = np.flatnonzero(done)
        if len(rows):
            want, ref, _ = orc.env_reset_rows(7, rows, episodes[rows], task, envs[0].ref_path.path_list, V)
            assert np.array_equal(o_g[rows].view(np.int32), want.view(np.int32)), (t, rows[:5])
            assert (envs[0].ref_indexes.cpu().numpy()[rows] == ref).all()
        assert 'reward_info' not in outs[0][3]
    assert n_done > 50
    assert (envs[0]._bufs['episode'].cpu().numpy() == episodes + 1).all()


def test_actions_read_in_place_from_a_cuda_tensor(e2e):
    """A policy's own CUDA action tensor passed again and again: after a few steps the env reads it in place
    (a graph captured for its address, no copy kernel); host arrays and other tensors keep going through the
    action buffer.  All three ways give identical steps."""
    task, B, V = 'right', 2048, 5
    envs = [e2e.CrossroadEnd2end(task, num_envs=B, veh_num=V, auto_reset=True, use_graph=True, reward_info=False)
            for _ in range(3)]
    for env in envs:
        env.seed(3)
        env.reset()
    rng = np.random.default_rng(5)
    mine = torch.empty((B, 2), device='cuda')                    # env 0: the same tensor every step
    for t in range(12):
        act = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
        mine.copy_(torch.as_tensor(act))
        envs[1].action_buffer.copy_(torch.as_tensor(act))         # env 1: the env's own buffer
        outs = [envs[0].step(mine), envs[1].step(envs[1].action_buffer), envs[2].step(act)]   # env 2: host array
        for o in outs[1:]:
            assert torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) and torch.equal(outs[0][2], o[2]), t
    keys = {k[1] for k in envs[0]._graphs}
    assert mine.data_ptr() in keys                               # the dedicated graphs exist ...
    assert {k[1] for k in envs[1]._graphs} == {None}            # (env 2's staging copies may recycle one address)
    # ... and a tensor seen once does not trigger a capture
    n = len(envs[0]._graphs)
    envs[0].step(torch.zeros((B, 2), device='cuda'))
    assert len(envs[0]._graphs) == n


@pytest.mark.parametrize('task', TASKS)
def test_set_traj_reprojects_tracking(e2e, task):
    """The reference's decision pattern `env.set_traj(path); env._get_obs()` (hier_decision.py:115-124):
    tracking columns re-projected onto the new path, and the next step keeps following it."""
    from env_build_b200.dynamics_and_models import ReferencePath
    env = e2e.CrossroadEnd2end(task)
    env.seed(11)
    obs0 = env.reset().copy()
    paths = env.ref_path.path_list
    for k in range(3):
        traj = ReferencePath(task, k)
        env.set_traj(traj)
        got = env._get_obs()
        rp = orc.ReferencePath(task, k, path_list=paths)
        want = rp.tracking_error_vector(obs0[3:4], obs0[4:5], obs0[5:6], obs0[0:1], 0)[0]
        assert np.array_equal(got[6:9].view(np.int32), want.view(np.int32)), k
        assert np.array_equal(got[:6], obs0[:6]) and np.array_equal(got[9:], obs0[9:])
        assert int(env.ref_indexes[0]) == k
    act = np.array([0.1, 0.2], np.float32)
    nxt, _, _, info = env.step(act)
    want_obs, _, _, _ = orc.gym_env_step(np.concatenate([obs0[:6], want, obs0[9:]])[None], act[None], task,
                                         np.array([2], np.int32), paths, orc.VEHICLE_MODE_LIST[task])
    assert info['ref_index'] == 2
    assert np.allclose(nxt[:6], want_obs[0, :6], rtol=1e-5, atol=1e-5)
    assert np.allclose(nxt[6:9], want_obs[0, 6:9], rtol=1e-5, atol=2e-5)


def test_assigned_obs_is_adopted(e2e):
    """env.obs / env.ref_indexes are plain attributes in the reference; assigning them must take effect."""
    from env_build_b200 import synthetic as syn
    rng = np.random.default_rng(1)
    task, B, V = 'left', 512, 8
    env = e2e.CrossroadEnd2end(task, num_envs=B)
    env.seed(2)
    env.reset()
    ref = syn.make_ref_indexes(rng, B)
    obs = syn.make_obs(rng, B, task, V, env.ref_path.path_list, ref)
    act = syn.make_actions(rng, 1, B)[0]
    env.obs = env.env_model._adopt(obs)
    env.ref_indexes = torch.as_tensor(ref, device='cuda')
    got = env.step(act)[0].numpy()
    want = orc.gym_env_step(obs, act, task, ref, env.ref_path.path_list, orc.VEHICLE_MODE_LIST[task])[0]
    assert np.allclose(got[:, :6], want[:, :6], rtol=1e-5, atol=1e-5)
