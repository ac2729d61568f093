#!/usr/bin/env python
"""Only bench.py's `sharded_from_rank0` leg, for NCCL tuning runs under torchrun:
    NCCL_MIN_P2P_NCHANNELS=8 python -m torch.distributed.run --nproc-per-node 8 ... tools/scatter_probe.py [rows per GPU]
Prints ms per rollout, env-steps/s and the source rank's egress rate."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                      # noqa: E402
from env_build_b200.dynamics_and_models import EnvironmentModel   # noqa: E402
from env_build_b200.parallel import ShardedRollout                # noqa: E402
from env_build_b200.rollout import RolloutGraph                   # noqa: E402

world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
H, V, D = bench.H, bench.V, bench.D
model = EnvironmentModel(bench.TASK, 0, mode='training', veh_mode_list=bench.mode_list())
sr = ShardedRollout(lambda b: RolloutGraph(model, b, V, H), world * B, D, H, dev, slots=2)
staged = None
if rank == 0:
    _, o, r, t = bench.make_inputs(B, 1)
    staged = sr.stage(np.tile(o, (world, 1)), np.tile(r, world), np.tile(t, (1, world, 1)))
for g in sr.runners:
    g.run()
comm = torch.cuda.Stream()
ready = [torch.cuda.Event() for _ in range(2)]
done = [torch.cuda.Event() for _ in range(2)]
for e in done:
    e.record(torch.cuda.current_stream())


def issue(slot):
    with torch.cuda.stream(comm):
        comm.wait_event(done[slot])
        sr.scatter_staged(staged, slot=slot)
        ready[slot].record(comm)


def rollouts(n, i0):
    main = torch.cuda.current_stream()
    for i in range(i0, i0 + n):
        slot = i % 2
        if i == 0:
            issue(0)
        issue(1 - slot)
        main.wait_event(ready[slot])
        sr.run(slot=slot)
        done[slot].record(main)
        sr.gather_returns(slot=slot)


rollouts(3, 0)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
rollouts(10, 3)
torch.cuda.current_stream().wait_stream(comm)
b.record()
torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b) / 10], device=dev, dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    nbytes = world * sr.runner.inbox.numel() * 4
    keys = [k for k in os.environ if k.startswith('NCCL_')]
    print('%s: %.3f ms per rollout, %.3g env-steps/s, source egress %.0f GB/s' % (
        ' '.join('%s=%s' % (k, os.environ[k]) for k in keys) or 'NCCL defaults', float(ms), world * B * H / (float(ms) * 1e-3),
        nbytes * (world - 1) / world / (float(ms) * 1e-3) / 1e9))
dist.destroy_process_group()
