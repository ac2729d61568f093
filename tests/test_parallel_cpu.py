"""Row sharding over ranks (env_build_b200/parallel.py) with world_size-2 gloo on CPU.
The compute inside each rank is the ORACLE here (tests may use it as a stand-in); what is under
test is the host-side partitioning, the ragged scatter/gather and shard invariance."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from env_build_b200 import parallel as par

H, V, TASK = 3, 5, 'right'


def test_shard_bounds_partition():
    for B in (0, 1, 7, 64, 65536, 1000003):
        for W in (1, 2, 3, 4, 8):
            cuts = [par.shard_bounds(B, W, r) for r in range(W)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(W - 1))
            sizes = par.shard_sizes(B, W)
            assert sum(sizes) == B and max(sizes) - min(sizes) <= 1


class OracleRunner(object):
    """Stand-in with RolloutGraph's interface (load / run / out5), CPU only."""

    def __init__(self, b):
        from oracle import crossroad_oracle as orc
        self.m = orc.EnvironmentModel(TASK, 0, mode='training')
        self.b = b

    def load(self, obs, ref, tape):
        self.obs, self.ref, self.tape = obs.numpy(), ref.numpy(), tape.contiguous().numpy()

    def run(self):
        self.m.reset(self.obs, self.ref)
        outs = []
        for t in range(self.tape.shape[0]):
            res = self.m.rollout_out(self.tape[t])
            outs.append(np.stack(res[1:]))
        self.out5 = torch.from_numpy(np.stack(outs))


def _inputs(B):
    from env_build_b200 import synthetic as syn
    from oracle import crossroad_oracle as orc
    rng = np.random.default_rng(42)
    paths = orc.construct_ref_paths(TASK)[0]
    ref = syn.make_ref_indexes(rng, B, out_of_range_frac=0.05)
    obs = syn.make_obs(rng, B, TASK, V, paths, ref)
    tape = syn.make_actions(rng, H, B)
    return obs, ref, tape


def _worker(rank, world, port, B, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        obs, ref, tape = _inputs(B)
        sr = par.ShardedRollout(OracleRunner, B, obs.shape[1], H, torch.device('cpu'))
        assert (sr.lo, sr.hi) == par.shard_bounds(B, world, rank)
        if rank == 0:
            sr.scatter(torch.from_numpy(obs), torch.from_numpy(ref), torch.from_numpy(tape))
        else:
            sr.scatter()
        assert sr.runner.obs.shape == (sr.hi - sr.lo, obs.shape[1])
        assert np.array_equal(sr.runner.obs, obs[sr.lo:sr.hi]) and np.array_equal(sr.runner.ref, ref[sr.lo:sr.hi])
        assert np.array_equal(sr.runner.tape, tape[:, sr.lo:sr.hi])
        sr.run()
        ret = sr.gather_returns()
        if rank == 0:
            np.save(out_path, ret.numpy())
        else:
            assert ret is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('B', [37, 64])
def test_sharded_rollout_gloo_world2(tmp_path, B):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'ret.npy')
    mp.spawn(_worker, args=(2, port, B, out), nprocs=2, join=True)
    got = np.load(out)
    # single-process result on the whole batch: shard invariance must be bit-exact
    obs, ref, tape = _inputs(B)
    r = OracleRunner(B)
    r.load(torch.from_numpy(obs), torch.from_numpy(ref), torch.from_numpy(tape))
    r.run()
    want = r.out5.sum(0).t().numpy()
    assert got.shape == (B, 5)
    assert np.array_equal(got.view(np.int32), want.view(np.int32))


class StaticRunner(OracleRunner):
    """OracleRunner with RolloutGraph's static buffers (obs0 padded rows, ref, tape), so that the
    staged scatter can write straight into them."""

    def __init__(self, b):
        from env_build_b200.dynamics_and_models import padded_rows
        OracleRunner.__init__(self, b)
        self.B = b
        self.model = type('M', (), dict(_veh_off=9))()
        probe = padded_rows(1, 9 + 4 * V, 9, torch.device('cpu'))
        ld, front = probe.stride(0), probe.storage_offset()
        n_obs = b * ld + 16
        self.inbox = torch.zeros(n_obs + H * b * 2 + b, dtype=torch.float32)
        self.obs0 = self.inbox[:n_obs].as_strided((b, 9 + 4 * V), (ld, 1), front)
        self.tape = self.inbox[n_obs:n_obs + H * b * 2].view(H, b, 2)
        self.ref = self.inbox[n_obs + H * b * 2:].view(torch.int32)

    def run(self):
        self.load(self.obs0.clone(), self.ref, self.tape)
        OracleRunner.run(self)


def _worker_staged(rank, world, port, B, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        obs, ref, tape = _inputs(B)
        sr = par.ShardedRollout(StaticRunner, B, obs.shape[1], H, torch.device('cpu'), slots=2)
        staged = sr.stage(obs, ref, tape) if rank == 0 else None
        for slot in (0, 1):
            sr.scatter_staged(staged, slot=slot)
            r = sr.runners[slot]
            assert np.array_equal(r.obs0.numpy(), obs[sr.lo:sr.hi]) and np.array_equal(r.ref.numpy(), ref[sr.lo:sr.hi])
            assert np.array_equal(r.tape.numpy(), tape[:, sr.lo:sr.hi])
        sr.run(slot=1)
        ret = sr.gather_returns(slot=1)
        if rank == 0:
            np.save(out_path, ret.numpy())
    finally:
        dist.destroy_process_group()


def test_staged_scatter_gloo_world2(tmp_path):
    """The no-copy exchange: views of the source's staged batch land in the ranks' static buffers."""
    B = 64
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'ret.npy')
    mp.spawn(_worker_staged, args=(2, port, B, out), nprocs=2, join=True)
    got = np.load(out)
    obs, ref, tape = _inputs(B)
    r = OracleRunner(B)
    r.load(torch.from_numpy(obs), torch.from_numpy(ref), torch.from_numpy(tape))
    r.run()
    assert np.array_equal(got.view(np.int32), r.out5.sum(0).t().numpy().view(np.int32))


def test_scatter_without_ref_or_tape():
    """mode != 'training' has no per-row path indexes, a closed-loop runner no tape (world size 1 here)."""
    class R(object):
        def __init__(self, b):
            pass

        def load(self, obs, ref, tape):
            self.got = (obs, ref, tape)
    sr = par.ShardedRollout(R, 8, 5, 3, torch.device('cpu'))
    sr.scatter(torch.zeros((8, 5)), None, None, has_ref=False, has_tape=False)
    assert sr.runner.got[0].shape == (8, 5) and sr.runner.got[1] is None and sr.runner.got[2] is None
