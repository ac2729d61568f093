"""Differentiable float64 restatement of EnvironmentModel.rollout_out in PyTorch (CPU).

TEST INFRASTRUCTURE -- NOT PRODUCT CODE (same rules as crossroad_oracle.py).  It exists to give
the backward kernel (ce2e_rollout_step_backward) a gradient oracle: torch.autograd differentiates
this forward exactly the way TensorFlow differentiates the reference's (same op graph:
tf.stop_gradient on the vehicle columns DM:195/331/402, integer argmin + gather DM:712-731,
tf.where selections, tf.clip_by_value).  The forward itself is checked against the NumPy oracle.
Citations: DM = reference dynamics_and_models.py.
"""
import math

import numpy as np
import torch

from . import crossroad_oracle as orc

D2R = math.pi / 180.0


def _circles(x, y, phi_deg):
    ang = phi_deg * D2R
    lws = (orc.L - orc.W) / 2.
    c, s = torch.cos(ang), torch.sin(ang)
    return (x + lws * c, y + lws * s), (x - lws * c, y - lws * s)


def _hinge(d, t):
    return torch.where(d - t < 0, (d - t) ** 2, torch.zeros_like(d))


def _road(task, p, real):
    px, py = p
    half, lw, lw2, lw3 = 25.0, 3.75, 7.5, 11.25
    z = torch.zeros_like(px)
    w = torch.where
    if task == 'left':
        third = (px < -half) if real else (px < 0)
        return (w((py < -half) & (px < 1), (px - 1) ** 2, z) + w((py < -half) & (lw - px < 1), (lw - px - 1) ** 2, z) +
                w(third & (lw3 - py < 1), (lw3 - py - 1) ** 2, z) + w((px < -half) & (py < 1), (py - 1) ** 2, z))
    if task == 'straight':
        return (w((py < -half) & (px - lw < 1), (px - lw - 1) ** 2, z) + w((py < -half) & (lw2 - px < 1), (lw2 - px - 1) ** 2, z) +
                w((py > half) & (lw3 - px < 1), (lw3 - px - 1) ** 2, z) + w((py > half) & (px < 1), (px - 1) ** 2, z))
    return (w((py < -half) & (px - lw2 < 1), (px - lw2 - 1) ** 2, z) + w((py < -half) & (lw3 - px < 1), (lw3 - px - 1) ** 2, z) +
            w((px > half) & (-py < 1), (-py - 1) ** 2, z) + w((px > half) & (py + lw3 < 1), (py + lw3 - 1) ** 2, z))


def rollout_out(obs, act_norm, task, ref_indexes, path_list, mode_list, num_future_data=0):
    """obs [B,D], act_norm [B,2] float64 torch tensors (requires_grad as wanted).  Returns
    (next_obs, rewards, punish_train, punish_real, veh2veh4real, veh2road4real)."""
    n = num_future_data
    ntr = 3 * (n + 1)
    ego, trk, veh = obs[:, :6], obs[:, 6:6 + ntr], obs[:, 6 + ntr:].detach()          # DM:189-195
    a = torch.clamp(act_norm, -1.05, 1.05)                                            # DM:129
    steer, a_x = 0.4 * a[:, 0], 2.25 * a[:, 1] - 0.75
    vx, vy, r, x, y, phi = (ego[:, i] for i in range(6))
    rewards = (0.05 * -(trk[:, 2] ** 2) + 0.8 * -(trk[:, 0] ** 2) + 30 * -((trk[:, 1] * D2R) ** 2) +
               0.02 * -(r ** 2) + 5 * -(steer ** 2) + 0.05 * -(a_x ** 2))              # DM:198-207, 297
    ef, er = _circles(x, y, phi)
    v2v_tr = torch.zeros_like(x)
    v2v_re = torch.zeros_like(x)
    for j in range(veh.shape[1] // 4):                                                # DM:218-229
        vf, vr = _circles(veh[:, 4 * j], veh[:, 4 * j + 1], veh[:, 4 * j + 3])
        for ep in (ef, er):
            for vp in (vf, vr):
                d = torch.sqrt((ep[0] - vp[0]) ** 2 + (ep[1] - vp[1]) ** 2)
                v2v_tr = v2v_tr + _hinge(d, 3.5)
                v2v_re = v2v_re + _hinge(d, 2.5)
    v2r_tr = _road(task, ef, False) + _road(task, er, False)                          # DM:231-295
    v2r_re = _road(task, ef, True) + _road(task, er, True)
    # ---- next obs (DM:322-358)
    p = orc.VEHICLE_PARAMS
    tau, m, Iz, Cf, Cr, la, lb = 0.1, p['mass'], p['I_z'], p['C_f'], p['C_r'], p['a'], p['b']
    ph = phi * D2R
    nvx = vx + tau * (a_x + vy * r)                                                   # DM:73-81
    nvy = (m * vy * vx + tau * (la * Cf - lb * Cr) * r - tau * Cf * steer * vx - tau * m * vx ** 2 * r) / \
        (m * vx - tau * (Cf + Cr))
    nr = (-Iz * r * vx - tau * (la * Cf - lb * Cr) * vy + tau * la * Cf * steer * vx) / \
        (tau * (la ** 2 * Cf + lb ** 2 * Cr) - Iz * vx)
    nx = x + tau * (vx * torch.cos(ph) - vy * torch.sin(ph))
    ny = y + tau * (vx * torch.sin(ph) + vy * torch.cos(ph))
    nphi = (ph + tau * r) * 180 / math.pi
    nvx = torch.clamp(nvx, 0., 35.)                                                   # DM:390
    ref_indexes = np.asarray(ref_indexes)
    B = obs.shape[0]
    trk_next = torch.zeros((B, ntr), dtype=obs.dtype)
    for pi, path in enumerate(path_list):                                             # DM:340-353
        msk = ref_indexes == pi
        if not msk.any():
            continue
        rp = orc.ReferencePath(task, pi, path_list=path_list)
        ex, ey, ephi, ev = nx[msk], ny[msk], nphi[msk], nvx[msk]
        idx, pts = rp.find_closest_point(ex.detach().numpy().astype(np.float32), ey.detach().numpy().astype(np.float32))
        rx, ry, rphi = (torch.from_numpy(np.asarray(q, dtype=np.float64)) for q in pts)
        half = 25.0
        if task == 'left':                                                            # DM:736-752
            dl = torch.sqrt((ex + half) ** 2 + (ey + half) ** 2) - torch.sqrt((rx + half) ** 2 + (ry + half) ** 2)
            dl = torch.where(ey < -half, ex - rx, dl)
            dl = torch.where(ex < -half, ey - ry, dl)
        elif task == 'straight':
            dl = ex - rx
        else:
            dl = -(torch.sqrt((ex - half) ** 2 + (ey + half) ** 2) - torch.sqrt((rx - half) ** 2 + (ry + half) ** 2))
            dl = torch.where(ey < -half, ex - rx, dl)
            dl = torch.where(ex > half, -(ey - ry), dl)

        def wrap(d):
            d = torch.where(d > 180., d - 360., d)
            return torch.where(d < -180., d + 360., d)
        cols = [-dl, wrap(ephi - rphi), ev - orc.EXPECTED_V]
        for fp in rp.future_n_data(idx, n):                                           # DM:763-768
            fx, fy, fphi = (torch.from_numpy(np.asarray(q, dtype=np.float64)) for q in fp)
            cols += [fx - ex, fy - ey, wrap(ephi - fphi)]
        trk_next[msk] = torch.stack(cols, 1)
    veh_next = torch.from_numpy(orc.veh_predict(veh.numpy().astype(np.float32), mode_list).astype(np.float64))
    next_obs = torch.cat([torch.stack([nvx, nvy, nr, nx, ny, nphi], 1), trk_next, veh_next], 1)
    return next_obs, rewards, v2v_tr + v2r_tr, v2v_re + v2r_re, v2v_re, v2r_re
