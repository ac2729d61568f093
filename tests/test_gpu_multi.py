"""Row sharding over GPUs with NCCL (env_build_b200/parallel.py).  Needs >= 2 GPUs; skipped otherwise.
Run on a multi-GPU box with:  python -m pytest tests/test_gpu_multi.py -m gpu"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, B, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from env_build_b200 import synthetic as syn
        from env_build_b200.dynamics_and_models import EnvironmentModel
        from env_build_b200.parallel import ShardedRollout
        from env_build_b200.rollout import RolloutGraph
        task, V, H = 'straight', 9, 5
        model = EnvironmentModel(task, mode='training')
        sr = ShardedRollout(lambda b: RolloutGraph(model, b, V, H), B, 45, H, dev)
        if rank == 0:
            rng = np.random.default_rng(11)
            ref = syn.make_ref_indexes(rng, B)
            obs = syn.make_obs(rng, B, task, V, model.ref_path.path_list, ref)
            tape = syn.make_actions(rng, H, B)
            sr.scatter(torch.from_numpy(obs).to(dev), torch.from_numpy(ref).to(dev), torch.from_numpy(tape).to(dev))
        else:
            sr.scatter()
        sr.run()
        ret = sr.gather_returns()
        if rank == 0:
            # single-GPU result on the whole batch: shard invariance must be bit-exact
            one = RolloutGraph(model, B, V, H)
            one.load(obs, ref, tape)
            one.run()
            want = one.out5.sum(0).t().contiguous()
            np.save(out_path, np.stack([ret.cpu().numpy(), want.cpu().numpy()]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('B', [4099])
def test_sharded_rollout_nccl(tmp_path, B):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    world = 2
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'ret.npy')
    mp.spawn(_worker, args=(world, port, B, out), nprocs=world, join=True)
    got, want = np.load(out)
    assert got.shape == (B, 5) and np.array_equal(got.view(np.int32), want.view(np.int32))


def _worker_staged(rank, world, port, B, exchange, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from env_build_b200 import synthetic as syn
        from env_build_b200.dynamics_and_models import EnvironmentModel
        from env_build_b200.parallel import ShardedRollout
        from env_build_b200.rollout import RolloutGraph
        task, V, H = 'left', 32, 4
        from env_build_b200.endtoend_env_utils import VEHICLE_MODE_LIST
        model = EnvironmentModel(task, mode='training', veh_mode_list=syn.tiled_mode_list(VEHICLE_MODE_LIST[task], V))
        sr = ShardedRollout(lambda b: RolloutGraph(model, b, V, H), B, 9 + 4 * V, H, dev, slots=2, exchange=exchange)
        rets, staged = [], None
        for k in range(2):                                  # two batches through the two buffer sets
            if rank == 0:
                rng = np.random.default_rng(21 + k)
                ref = syn.make_ref_indexes(rng, B)
                obs = syn.make_obs(rng, B, task, V, model.ref_path.path_list, ref)
                tape = syn.make_actions(rng, H, B)
                staged = sr.stage(obs, ref, tape)
            sr.scatter_staged(staged, slot=k)
            sr.run(slot=k)
            ret = sr.gather_returns(slot=k)
            if rank == 0:
                one = RolloutGraph(model, B, V, H)
                one.load(obs, ref, tape)
                one.run()
                rets.append(np.stack([ret.cpu().numpy(), one.out5.sum(0).t().contiguous().cpu().numpy()]))
        if rank == 0:
            np.save(out_path, np.stack(rets))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('exchange', ['nccl', 'peer'])
def test_staged_exchange_two_gpus(tmp_path, exchange):
    """Staged batch on rank 0 -> the ranks' inboxes by NCCL scatter or by copy-engine pulls over the NVLink
    peer mapping; both bit-identical to one GPU rolling out the whole batch."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    B = 4096
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'ret.npy')
    mp.spawn(_worker_staged, args=(2, port, B, exchange, out), nprocs=2, join=True)
    for got, want in np.load(out):
        assert got.shape == (B, 5) and np.array_equal(got.view(np.int32), want.view(np.int32))


def test_one_process_two_devices():
    """One process driving two GPUs: the same EnvironmentModel / ReferencePath objects keep one table
    handle per device, inputs follow the current device; results are bit-identical on both."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    from env_build_b200 import synthetic as syn
    from env_build_b200.dynamics_and_models import EnvironmentModel
    rng = np.random.default_rng(12)
    task, B, V = 'left', 3001, 8
    model = EnvironmentModel(task, mode='training')
    ref = syn.make_ref_indexes(rng, B)
    obs = syn.make_obs(rng, B, task, V, model.ref_path.path_list, ref)
    tape = syn.make_actions(rng, 3, B)
    outs = []
    for d in (0, 1, 0):
        with torch.cuda.device(d):
            model.reset(obs, ref)
            for t in range(3):
                res = model.rollout_out(tape[t])
            assert res[0].device.index == d
            outs.append([r.numpy() for r in res])
    assert len(model.ref_path._handles) == 2
    for a, b, c in zip(*outs):
        assert np.array_equal(a.view(np.int32), b.view(np.int32)) and np.array_equal(a.view(np.int32), c.view(np.int32))
