"""Host-side logic and the C-ABI surface, no GPU needed: path-table construction against the
golden tables, the library builds / loads / exports every symbol include/ce2e.h declares,
argument validation that happens before any CUDA call, the padded row layout."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, TASKS
from env_build_b200 import _lib
from env_build_b200 import dynamics_and_models as dm
from env_build_b200 import endtoend_env_utils as eu


@pytest.fixture(scope='module')
def lib():
    if _lib.needs_build():
        _lib.build()
    return _lib.load()


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'ce2e.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ce2e_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(lib):
    names = declared_functions()
    assert len(names) >= 15
    assert sorted(_lib.SIGNATURES) == names, 'ctypes table and include/ce2e.h disagree'
    for n in names:
        assert hasattr(lib, n), n
    assert lib.ce2e_version() == 100


def test_header_cites_reference():
    text = open(os.path.join(ROOT, 'include', 'ce2e.h')).read()
    for fn in declared_functions():
        if fn in ('ce2e_version', 'ce2e_last_error', 'ce2e_launch_count', 'ce2e_paths_destroy'):
            continue
        i = text.index(fn + '(')
        doc = text[text.rfind('/*', 0, i):i]
        assert re.search(r'(DM|E2E|EU):\d+', doc), '%s: no reference citation' % fn


def test_argument_validation_without_gpu(lib):
    """Checks that run before the first CUDA call return the documented codes."""
    h = ctypes.c_void_p()
    lens = (ctypes.c_int32 * 1)(100)
    arr = (ctypes.c_void_p * 1)(None)
    assert lib.ce2e_paths_create(7, 1, lens, arr, arr, arr, ctypes.byref(h)) == -3          # CE2E_ERR_TASK
    assert b'task' in lib.ce2e_last_error()
    assert lib.ce2e_paths_create(0, 9, lens, arr, arr, arr, ctypes.byref(h)) == -2          # CE2E_ERR_SHAPE
    assert lib.ce2e_paths_create(0, 1, lens, arr, arr, arr, ctypes.byref(h)) == -1          # CE2E_ERR_NULL
    assert lib.ce2e_action_transform(None, None, 4, None) == -1
    assert lib.ce2e_dynamics_step(None, 6, None, 0.1, None, 6, None, 0, 4, None) == -1
    assert lib.ce2e_compute_rewards(5, None, 41, None, 8, 0, None, None, 4, None) == -3
    assert lib.ce2e_compute_rewards(0, None, 41, None, 8, 0, None, None, -1, None) == -2
    assert lib.ce2e_rollout_step(None, 0, None, None, 0, None, None, 0, 0, 0, None, 0, None, None, 1, None) == -1
    with pytest.raises(ValueError):
        _lib.check(-2)
    with pytest.raises(_lib.Ce2eError):
        _lib.check(-5)


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    m = dm.EnvironmentModel('left', mode='selecting')
    with pytest.raises(RuntimeError):
        m.add_traj(np.zeros((2, 41), np.float32), 0)
    with pytest.raises(RuntimeError):
        dm.VehicleDynamics().f_xu(np.zeros((2, 6), np.float32), np.zeros((2, 2), np.float32), 0.1)


@pytest.mark.parametrize('task', TASKS)
def test_path_tables_match_reference(task, golden_task):
    g = golden_task(task)
    paths, len_list, ctrl = dm.build_path_tables(task)
    assert (np.array(len_list) == g['path_len_list']).all()
    assert np.array_equal(np.array(ctrl, dtype=np.float64), g['control_points'])
    for i, p in enumerate(paths):
        for c, name in enumerate('xy'):
            a, b = p[c], g['path%d_%s' % (i, name)]
            assert a.dtype == np.float32 and a.shape == b.shape and (a.view(np.int32) == b.view(np.int32)).all()
        a, b = p[2], g['path%d_phi' % i]
        # the reference's heading is NumPy's fp32 arctan2 (SIMD build dependent, <= 4 ulp);
        # the product (like the oracle) rounds the float64 atan2
        ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
        ulp[(a == 0) & (b == 0)] = 0
        assert ulp.max() <= 4
    rp = dm.ReferencePath(task, 2)
    assert rp.path is rp.path_list[2] and len(rp.path_list) == 3
    rp.set_path(0)
    assert rp.ref_index == 0 and rp.path is rp.path_list[0]
    with pytest.raises(IndexError):
        rp.set_path(5)


def test_constants_and_mode_tables(golden_common):
    c = golden_common
    assert (eu.L, eu.W, eu.LANE_WIDTH, eu.LANE_NUMBER, eu.CROSSROAD_SIZE, eu.EXPECTED_V) == \
        tuple(float(c['const_' + k]) for k in ('L', 'W', 'LANE_WIDTH', 'LANE_NUMBER', 'CROSSROAD_SIZE', 'EXPECTED_V'))
    for task in TASKS:
        assert list(c['mode_list_' + task]) == eu.VEHICLE_MODE_LIST[task]
        assert eu.VEH_NUM[task] == len(eu.VEHICLE_MODE_LIST[task])
        got = np.array([eu.judge_feasible(x, y, task) for x, y in c['jf_xy']])
        assert (got == c['jf_' + task]).all()
    assert [eu.deal_with_phi(float(p)) for p in c['dwp_in']] == list(c['dwp_out'])
    vd = dm.VehicleDynamics()
    for k, v in vd.vehicle_params.items():
        assert float(c['vp_' + k]) == v
    assert [eu.turn_class(m) for m in ('dl', 'rd', 'ur', 'lu', 'dr', 'ru', 'ul', 'ld', 'du', 'lr')] == \
        [1, 1, 1, 1, -1, -1, -1, -1, 0, 0]


@pytest.mark.parametrize('D,veh_off', [(137, 9), (41, 9), (45, 9), (29, 9), (71, 39), (50, 18), (9, 9)])
def test_padded_rows_alignment(D, veh_off):
    t = dm.padded_rows(5, D, veh_off, device='cpu')
    assert t.shape == (5, D) and t.stride(1) == 1 and t.stride(0) % 4 == 0
    assert (t.storage_offset() + veh_off) % 16 == 0                     # row 0: 64 B; every row: 16 B
    assert t.stride(0) - D >= (7 if veh_off == 9 else 0)                # room for the ego window box
    assert t.stride(0) - D < 11
    t.copy_(torch.arange(5 * D, dtype=torch.float32).reshape(5, D))
    assert t[4, D - 1] == 5 * D - 1


def test_div_const_exact():
    """The kernels replace x/c by q=x*rc; r=fma(-c,q,x); q+=r*rc (ce2e_device.cuh div_const).
    tests/tools/verify_divc.c checks bit-equality with IEEE division (exhaustive run recorded in
    DESIGN.md); here a 2^24-value stride sample per divisor."""
    import subprocess
    import tempfile
    src = os.path.join(ROOT, 'tests', 'tools', 'verify_divc.c')
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, 'verify_divc')
        subprocess.run(['gcc', '-O2', '-ffp-contract=off', '-o', exe, src, '-lm'], check=True)
        out = subprocess.run([exe, '257'], check=True, capture_output=True, text=True).stdout
    assert 'mismatches=0' in out and 'FAIL' not in out, out


@pytest.mark.parametrize('task', TASKS)
def test_candidate_grid_contains_bruteforce_argmin(lib, task):
    """ce2e_grid.h: for every query point the fp32 brute-force first-argmin (the oracle's
    find_closest_point, DM:702-715) must lie inside the [lo, hi] range of the point's cell.
    Points: uniform over the map, concentrated near the paths, and on / next to cell edges."""
    from oracle import crossroad_oracle as orc
    paths = dm.build_path_tables(task)[0]
    rng = np.random.default_rng(123)
    for pi, p in enumerate(paths):
        wx, wy = np.ascontiguousarray(p[0][::10]), np.ascontiguousarray(p[1][::10])
        spec = np.zeros(5, np.float32)
        assert lib.ce2e_grid_build_host(wx.ctypes.data, wy.ctypes.data, len(wx), spec.ctypes.data, None, 0) == 0
        x0, y0, inv_h, nx, ny = spec[0], spec[1], spec[2], int(spec[3]), int(spec[4])
        cells = np.zeros(nx * ny, np.uint32)
        assert lib.ce2e_grid_build_host(wx.ctypes.data, wy.ctypes.data, len(wx), spec.ctypes.data,
                                        cells.ctypes.data, cells.size) == 0
        n = 60000
        k = rng.integers(0, len(wx), n)
        near = np.stack([wx[k] + rng.normal(0, 1.5, n), wy[k] + rng.normal(0, 1.5, n)], 1)
        uni = np.stack([rng.uniform(x0 - 5, x0 + nx / inv_h + 5, n), rng.uniform(y0 - 5, y0 + ny / inv_h + 5, n)], 1)
        edge = np.stack([x0 + rng.integers(0, nx, n) / inv_h, y0 + rng.uniform(0, ny, n) / inv_h], 1)
        edge2 = np.stack([wx[k] + rng.normal(0, 1.0, n), y0 + rng.integers(0, ny, n) / inv_h], 1).astype(np.float32)
        edge2[:, 1] = np.nextafter(edge2[:, 1], np.float32(-1e9))                 # just below a cell edge
        onpt = np.stack([wx[k], wy[k]], 1)                                       # exactly on waypoints
        mid = np.stack([(wx[k] + wx[np.minimum(k + 1, len(wx) - 1)]) / 2,
                        (wy[k] + wy[np.minimum(k + 1, len(wx) - 1)]) / 2], 1)     # bisector points (ties)
        pts = np.concatenate([near, uni, edge, edge2, onpt, mid]).astype(np.float32)
        xs, ys = pts[:, 0], pts[:, 1]
        rp = orc.ReferencePath(task, pi, path_list=paths)
        want = rp.find_closest_point(xs, ys)[0] // 10
        # the device's cell computation, in fp32 (candidate_range in ce2e.cu)
        tx = (xs - np.float32(x0)) * np.float32(inv_h)
        ty = (ys - np.float32(y0)) * np.float32(inv_h)
        inside = (tx >= 0) & (ty >= 0) & (tx < np.float32(nx)) & (ty < np.float32(ny))
        c = cells[(ty[inside].astype(np.int32) * nx + tx[inside].astype(np.int32))]
        lo, hi = (c & 0xffff).astype(np.int64), (c >> 16).astype(np.int64)
        w = want[inside]
        assert inside.mean() > 0.5
        assert ((lo <= w) & (w <= hi)).all(), (task, pi, int(((lo > w) | (w > hi)).sum()))
        # near the path the ranges are a handful of waypoints (that is the point of the grid)
        d2 = ((xs[inside][:20000, None] - wx[None, :]) ** 2 + (ys[inside][:20000, None] - wy[None, :]) ** 2).min(1)
        assert (hi - lo + 1)[:20000][d2 < 9].max() <= 8


@pytest.mark.parametrize('kind', ['loop', 'zigzag', 'duplicates', 'random_walk', 'short'])
def test_candidate_grid_arbitrary_paths(lib, kind):
    """The C ABI accepts any path table (ce2e_paths_create), so the exactness of the candidate grid must
    not depend on the three built-in geometries: self-crossing loops, jittery zigzags, repeated
    waypoints (distance ties -> FIRST minimum), random walks, very short tables."""
    rng = np.random.default_rng({'loop': 1, 'zigzag': 2, 'duplicates': 3, 'random_walk': 4, 'short': 5}[kind])
    if kind == 'loop':                                 # a figure of eight, crossing itself
        t = np.linspace(0, 4 * np.pi, 420)
        wx, wy = 30 * np.sin(t), 20 * np.sin(2 * t)
    elif kind == 'zigzag':
        wx = np.linspace(-40, 40, 380)
        wy = 3 * np.sign(np.sin(np.arange(380) * 1.3)) + rng.normal(0, 0.05, 380)
    elif kind == 'duplicates':                         # every waypoint three times: ties everywhere
        t = np.repeat(np.linspace(0, 1, 130), 3)
        wx, wy = -25 + 50 * t, 10 * np.sin(6 * t)
    elif kind == 'random_walk':
        wx, wy = np.cumsum(rng.normal(0, 0.8, 400)), np.cumsum(rng.normal(0, 0.8, 400))
    else:
        wx, wy = np.array([0., 1., 2.]), np.array([0., 0.5, 0.])
    wx, wy = np.ascontiguousarray(wx, np.float32), np.ascontiguousarray(wy, np.float32)
    spec = np.zeros(5, np.float32)
    assert lib.ce2e_grid_build_host(wx.ctypes.data, wy.ctypes.data, len(wx), spec.ctypes.data, None, 0) == 0
    x0, y0, inv_h, nx, ny = spec[0], spec[1], spec[2], int(spec[3]), int(spec[4])
    cells = np.zeros(nx * ny, np.uint32)
    assert lib.ce2e_grid_build_host(wx.ctypes.data, wy.ctypes.data, len(wx), spec.ctypes.data, cells.ctypes.data,
                                    cells.size) == 0
    n = 40000
    k = rng.integers(0, len(wx), n)
    pts = np.concatenate([
        np.stack([wx[k] + rng.normal(0, 1.0, n), wy[k] + rng.normal(0, 1.0, n)], 1),
        np.stack([rng.uniform(x0, x0 + nx / inv_h, n), rng.uniform(y0, y0 + ny / inv_h, n)], 1),
        np.stack([wx[k], wy[k]], 1),
        np.stack([(wx[k] + wx[(k + 1) % len(wx)]) / 2, (wy[k] + wy[(k + 1) % len(wx)]) / 2], 1)]).astype(np.float32)
    xs, ys = pts[:, 0], pts[:, 1]
    d2 = np.square(xs[:, None] - wx[None, :]) + np.square(ys[:, None] - wy[None, :])      # fp32, DM:712
    want = d2.argmin(1)                                                                   # first minimum
    tx = (xs - np.float32(x0)) * np.float32(inv_h)
    ty = (ys - np.float32(y0)) * np.float32(inv_h)
    inside = (tx >= 0) & (ty >= 0) & (tx < np.float32(nx)) & (ty < np.float32(ny))
    c = cells[(ty[inside].astype(np.int32) * nx + tx[inside].astype(np.int32))]
    lo, hi = (c & 0xffff).astype(np.int64), (c >> 16).astype(np.int64)
    w = want[inside]
    assert inside.mean() > 0.5
    assert ((lo <= w) & (w <= hi)).all(), (kind, int(((lo > w) | (w > hi)).sum()))


def test_makefile_flags_match():
    """The Makefile builds the same library as env_build_b200._lib.build()."""
    from env_build_b200 import _lib
    mk = open(os.path.join(ROOT, 'Makefile')).read()
    line = [l for l in mk.splitlines() if l.startswith('NVCC_FLAGS')][0]
    assert line.split(':=')[1].split() == _lib.NVCC_FLAGS
