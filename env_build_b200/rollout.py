"""Graph-captured multi-step rollouts on static device buffers.

`EnvironmentModel.rollout_out` allocates its outputs per call like the reference does.  For
launch-bound inner loops (the reference's shield rollouts: hier_decision.py:89-97, H=5;
multi_ego.py:187-197, H=20; MPC horizon 25, mpc_ipopt.py:330) `RolloutGraph` keeps the
observation ping-pong buffers, the action tape and the per-step outputs resident and replays
the H fused `ce2e_rollout_step` launches as ONE CUDA graph.  Same kernels, same results.
"""
import ctypes

import torch

from . import _lib
from .dynamics_and_models import _ptr, _wrap, padded_rows, to_device


class RolloutGraph(object):
    """H open-loop steps of `model.rollout_out` for B rows with V vehicles each.

    load(obses, ref_indexes, tape)   copy inputs (host or device) into the static buffers
    run()                            replay the graph (H launches); results in
                                     .out5 [H,5,B] (rewards, punish_train, punish_real,
                                     veh2veh4real, veh2road4real per step) and .final_obs [B,D]

    Closed loop (the reference's shield pattern `actions = policy(obses); model.rollout_out(actions)`,
    hier_decision.py:93-96): pass `policy`, a callable mapping the current observations (a [B,D]
    CUDA tensor view, rows padded) to normalised actions [B,2] with capture-safe torch ops; it is
    captured into the same graph between the step launches and the actions it chose are left in
    .tape [H,B,2].
    """

    def __init__(self, model, B, V, H, use_graph=True, fused=False, policy=None):
        """fused=True: ONE launch of the horizon-fused kernel (ce2e_rollout_horizon: the tile's state
        stays on chip across the H steps) instead of H launches; same results, V <= 32, open loop only."""
        self.fused = bool(fused)
        self.policy = policy
        if policy is not None and self.fused:
            raise ValueError('the horizon-fused kernel needs the whole action tape in advance (no policy)')
        self.model, self.B, self.V, self.H = model, int(B), int(V), int(H)
        if len(model.veh_mode_list) != V:
            raise ValueError('the model predicts %d vehicles, rows hold %d' % (len(model.veh_mode_list), V))
        self.n = int(model.num_future_data)
        off = model._veh_off
        self.D = off + 4 * V
        dev = torch.device('cuda', torch.cuda.current_device())
        # the three inputs live in ONE allocation ("inbox": padded observation rows | action tape | path
        # indexes), so that a sharded caller can deliver all of them with a single collective
        probe = padded_rows(1, self.D, off, dev)
        ld, front = probe.stride(0), probe.storage_offset()
        n_obs = max(B, 1) * ld + 16
        self.inbox = torch.zeros(n_obs + H * B * 2 + B, dtype=torch.float32, device=dev)
        self.obs0 = self.inbox[:n_obs].as_strided((B, self.D), (ld, 1), front)
        self.tape = self.inbox[n_obs:n_obs + H * B * 2].view(H, B, 2)
        self.ref = self.inbox[n_obs + H * B * 2:].view(torch.int32)
        self.buf = [padded_rows(B, self.D, off, dev) for _ in range(2)]
        self.out5 = torch.zeros((H, 5, B), dtype=torch.float32, device=dev)
        self.final_obs = _wrap(self.buf[(H - 1) % 2])
        self._graph = None
        self._graph_key, self._graph_path, self._filled_index = None, None, None
        self.use_graph = use_graph
        self.launches_per_run = 0

    def load(self, obses, ref_indexes=None, tape=None):
        self.obs0.copy_(to_device(obses), non_blocking=True)
        if ref_indexes is not None:
            self.ref.copy_(to_device(ref_indexes, torch.int32).reshape(-1), non_blocking=True)
            self._filled_index = None
        if tape is not None:
            self.tape.copy_(to_device(tape), non_blocking=True)

    def _enqueue(self):
        m, lib = self.model, _lib.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        # the path index always travels through the device-side `ref` buffer (filled by run() when the model
        # is not in 'training' mode), so a captured graph follows later add_traj / set_path calls
        path_index, ref = 0, _ptr(self.ref)
        handle = m.ref_path.handle
        src = self.obs0
        ld = self.obs0.stride(0)
        if self.fused:
            dst = self.buf[(self.H - 1) % 2]
            _lib.check(lib.ce2e_rollout_horizon(handle, path_index, ref, _ptr(src), ld, _ptr(self.tape),
                                                ctypes.byref(m._turn), self.V, self.n, self.H, _ptr(dst), ld,
                                                _ptr(self.out5), self.B, stream))
            return
        for t in range(self.H):
            dst = self.buf[t % 2]
            if self.policy is not None:
                with torch.no_grad():
                    self.tape[t].copy_(self.policy(src))
            _lib.check(lib.ce2e_rollout_step(handle, path_index, ref, _ptr(src), ld, _ptr(self.tape[t]),
                                             ctypes.byref(m._turn), self.V, self.V, self.n, _ptr(dst), ld,
                                             _ptr(self.out5[t]), None, self.B, stream))
            src = dst

    def run(self):
        m = self.model
        if m.mode != 'training':
            idx = int(m.ref_path.ref_index)
            if idx != self._filled_index:
                self.ref.fill_(idx)
                self._filled_index = idx
        # a captured graph holds the table handle of the ReferencePath it was captured with
        key = (id(m.ref_path), m.ref_path.handle.value if hasattr(m.ref_path.handle, 'value') else int(m.ref_path.handle))
        if self._graph is not None and key != self._graph_key:
            self._graph = None
        self._graph_key, self._graph_path = key, m.ref_path
        if not self.use_graph:
            self._enqueue()
            self.launches_per_run = 1 if self.fused else self.H
            return
        if self._graph is None:
            self.model.ref_path.handle            # create the device tables outside the capture
            n0 = _lib.launch_count()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._enqueue()                   # warm-up launch outside the capture
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            n1 = _lib.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self.launches_per_run = _lib.launch_count() - n1
            assert self.launches_per_run == n1 - n0 == (1 if self.fused else self.H)
            self._graph = g
        self._graph.replay()
