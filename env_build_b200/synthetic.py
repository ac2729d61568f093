"""Seeded synthetic inputs for the model hot path.

Pure NumPy, no device code: the same generator feeds the golden-vector script,
the parity tests and bench.py, so every arm sees identical inputs.

Observation row layout (reference endtoend.py:300, dynamics_and_models.py:189-194):
    [v_x, v_y, r, x, y, phi_deg | d_y, d_phi_deg, d_v | (dx, dy, dphi) * n | (x, y, v, phi_deg) * V]
"""
import numpy as np

# start-index window of CrossroadEnd2end._reset_init_state (reference endtoend.py:473-478)
_RESET_SPAN = dict(left=900 + 500, straight=1200 + 500, right=420 + 500)

SEED_BASE = 20210310


def obs_dim(V, num_future_data=0):
    return 6 + 3 * (num_future_data + 1) + 4 * V


def make_ref_indexes(rng, B, n_paths=3, out_of_range_frac=0.0):
    """Per-row active path index for mode='training'; a fraction may be set to
    n_paths (no path matches -> zero tracking, reference dynamics_and_models.py:342-353)."""
    idx = rng.integers(0, n_paths, size=B).astype(np.int32)
    if out_of_range_frac > 0:
        m = rng.random(B) < out_of_range_frac
        idx[m] = n_paths
    return idx


def make_obs(rng, B, task, V, path_list, ref_indexes=0, num_future_data=0, edge_frac=0.05, near_frac=0.25):
    """[B, D] float32 observations.

    rng          np.random.Generator
    path_list    the task's 3 reference paths, each (x, y, phi_deg) float32 arrays
    ref_indexes  int or int array [B]: path each ego starts near (values >= len(path_list) use path 0)
    """
    n = num_future_data
    D = obs_dim(V, n)
    ref = np.broadcast_to(np.asarray(ref_indexes, dtype=np.int64), (B,)).copy()
    ref[(ref < 0) | (ref >= len(path_list))] = 0
    px = np.stack([p[0] for p in path_list])[:, :min(len(p[0]) for p in path_list)]
    py = np.stack([p[1][:px.shape[1]] for p in path_list])
    pphi = np.stack([p[2][:px.shape[1]] for p in path_list])
    span = min(_RESET_SPAN[task], px.shape[1] - 701)
    wp = rng.integers(700, 700 + span, size=B)
    obs = np.zeros((B, D), dtype=np.float32)
    v_x = rng.uniform(0, 8, B)
    obs[:, 0] = v_x
    obs[:, 1] = rng.uniform(-0.5, 0.5, B)
    obs[:, 2] = rng.uniform(-0.3, 0.3, B)
    obs[:, 3] = px[ref, wp] + rng.normal(0, 0.5, B)
    obs[:, 4] = py[ref, wp] + rng.normal(0, 0.5, B)
    obs[:, 5] = pphi[ref, wp] + rng.normal(0, 5.0, B)
    obs[:, 6] = rng.normal(0, 0.5, B)
    obs[:, 7] = rng.normal(0, 5.0, B)
    obs[:, 8] = obs[:, 0] - np.float32(8.)
    if n > 0:
        fut = obs[:, 9:9 + 3 * n].reshape(B, n, 3)
        fut[:, :, 0] = rng.normal(0, 8.0, (B, n))
        fut[:, :, 1] = rng.normal(0, 8.0, (B, n))
        fut[:, :, 2] = rng.normal(0, 10.0, (B, n))
    if V > 0:
        veh = obs[:, 9 + 3 * n:].reshape(B, V, 4)
        near = rng.random((B, V)) < near_frac
        vx = np.where(near, obs[:, 3:4] + rng.uniform(-8, 8, (B, V)), rng.uniform(-65, 65, (B, V)))
        vy = np.where(near, obs[:, 4:5] + rng.uniform(-8, 8, (B, V)), rng.uniform(-65, 65, (B, V)))
        veh[:, :, 0] = vx
        veh[:, :, 1] = vy
        veh[:, :, 2] = rng.uniform(0, 8, (B, V))
        veh[:, :, 3] = rng.choice(np.array([0., 90., 180., -90.]), size=(B, V)) + rng.normal(0, 10.0, (B, V))
    # edge rows: standstill, heading on the +-180 seam, ego off the lanes (road penalties fire)
    if edge_frac > 0 and B >= 8:
        e = rng.random(B)
        obs[e < edge_frac * 0.25, 0] = 0.0
        m = (e >= edge_frac * 0.25) & (e < edge_frac * 0.5)
        obs[m, 5] = np.where(rng.random(m.sum()) < 0.5, 180.0, -180.0)
        m = (e >= edge_frac * 0.5) & (e < edge_frac)
        obs[m, 3] += rng.uniform(-6, 6, m.sum())
        obs[m, 4] += rng.uniform(-6, 6, m.sum())
    return obs


def make_actions(rng, H, B):
    """Open-loop normalised action tape [H, B, 2] ~ U(-1, 1); a few values beyond
    the +-1.05 clip (reference dynamics_and_models.py:129)."""
    a = rng.uniform(-1, 1, (H, B, 2)).astype(np.float32)
    if B >= 8:
        m = rng.random((H, B)) < 0.02
        a[m, 0] = np.float32(1.3)
        m = rng.random((H, B)) < 0.02
        a[m, 1] = np.float32(-1.2)
    return a


def tiled_mode_list(native_list, V):
    """The task's vehicle-mode list tiled to V entries."""
    reps = -(-V // len(native_list))
    return (list(native_list) * reps)[:V]
