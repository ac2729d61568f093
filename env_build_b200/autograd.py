"""Differentiable `rollout_out`.

The reference's model-based trainer back-propagates through `EnvironmentModel.rollout_out`
(TensorFlow autodiff; vehicle columns under tf.stop_gradient, DM:195/331/402).  Here the forward
is the fused `ce2e_rollout_step` and the backward is `ce2e_rollout_step_backward`; both recompute
nothing on the host.  `EnvironmentModel.rollout_out` routes through this Function whenever the
actions or the stored observations require grad.
"""
import ctypes

import torch

from . import _lib


def _vp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class RolloutStep(torch.autograd.Function):
    """(obs [B,D] padded rows, act_norm [B,2]) -> (next_obs [B,D'], out5 [5,B])."""

    @staticmethod
    def forward(ctx, obs, act, model, path_index, ref, V_in, V_out):
        from .dynamics_and_models import padded_rows
        B = obs.shape[0]
        veh_off = model._veh_off
        obs_d, act_d = obs.detach(), act.detach().contiguous()
        nxt = padded_rows(B, veh_off + 4 * V_out, veh_off, obs.device)
        out5 = torch.empty((5, B), dtype=torch.float32, device=obs.device)
        scaled = torch.empty((B, 2), dtype=torch.float32, device=obs.device)
        ld = obs_d.stride(0) if B > 1 else max(obs_d.stride(0), obs_d.shape[1])
        _lib.check(_lib.load().ce2e_rollout_step(model.ref_path.handle, path_index, _vp(ref), _vp(obs_d), ld, _vp(act_d),
                                                 model._turn_ref, V_in, V_out, int(model.num_future_data), _vp(nxt),
                                                 nxt.stride(0) if B > 1 else max(nxt.stride(0), nxt.shape[1]),
                                                 _vp(out5), _vp(scaled), B, _stream()))
        ctx.model, ctx.path_index, ctx.V_in = model, path_index, V_in
        ctx.save_for_backward(obs_d, act_d, ref if ref is not None else torch.empty(0, device=obs.device))
        ctx.has_ref = ref is not None
        ctx.mark_non_differentiable(scaled)
        return nxt, out5, scaled

    @staticmethod
    def backward(ctx, g_next, g_out5, _g_scaled):
        obs, act, ref = ctx.saved_tensors
        model = ctx.model
        B, D = obs.shape
        n_cols = model._veh_off
        g_next = torch.zeros((B, n_cols), dtype=torch.float32, device=obs.device) if g_next is None else \
            g_next[:, :n_cols].to(torch.float32).contiguous()
        g_out5 = torch.zeros((5, B), dtype=torch.float32, device=obs.device) if g_out5 is None else \
            g_out5.to(torch.float32).contiguous()
        g_obs = torch.zeros((B, D), dtype=torch.float32, device=obs.device)      # vehicle columns: stop_gradient
        g_act = torch.empty((B, 2), dtype=torch.float32, device=obs.device)
        ld = obs.stride(0) if B > 1 else max(obs.stride(0), obs.shape[1])
        _lib.check(_lib.load().ce2e_rollout_step_backward(
            model.ref_path.handle, ctx.path_index, _vp(ref) if ctx.has_ref else None, _vp(obs), ld, _vp(act), ctx.V_in,
            int(model.num_future_data), _vp(g_next), n_cols, _vp(g_out5), _vp(g_obs), D, _vp(g_act), B, _stream()))
        return g_obs, g_act, None, None, None, None, None
