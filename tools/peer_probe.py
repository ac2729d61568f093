#!/usr/bin/env python
"""Copy-engine exchange probe: how fast can the ranks PULL their block out of rank 0's memory (or rank 0
PUSH it) with plain cudaMemcpyAsync over a peer mapping, no NCCL kernels on the SMs?
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/peer_probe.py [MB per rank]
The mapping is torch's symmetric memory (CUDA VMM handles); raw cudaIpcOpenMemHandle fails in this container."""
import os
import sys

import torch
import torch.distributed as dist

world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
MB = int(sys.argv[1]) if len(sys.argv) > 1 else 51
n = MB * (1 << 20) // 4                      # floats per rank block


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def report(tag, ms):
    say('%-34s %.3f ms  source egress %.0f GB/s' % (tag, ms, (world - 1) * n * 4 / (ms * 1e-3) / 1e9))


# ---- NCCL scatter of the same bytes, for scale
src = torch.arange(world * n, dtype=torch.float32, device=dev) if rank == 0 else None
dst = torch.empty(n, dtype=torch.float32, device=dev)
report('nccl scatter', timed(lambda: dist.scatter(dst, list(src.view(world, n).unbind(0)) if rank == 0 else None, src=0)))

# ---- torch symmetric memory
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(world * n, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    if rank == 0:
        t.copy_(src)
    torch.cuda.synchronize(); dist.barrier()
    root = hdl.get_buffer(0, (world, n), torch.float32)

    def pull():
        if rank != 0:
            dst.copy_(root[rank], non_blocking=True)
    report('symm pull (each rank its block)', timed(pull))
    torch.cuda.synchronize(); dist.barrier()
    ok = bool(rank == 0 or torch.equal(dst, torch.arange(rank * n, (rank + 1) * n, dtype=torch.float32, device=dev)))
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    say('symm pull content', 'ok' if int(flag) else 'WRONG')

    peers = [hdl.get_buffer(r, (world, n), torch.float32) for r in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]

    def push():
        if rank == 0:
            cur = torch.cuda.current_stream()
            for r in range(1, world):
                streams[r].wait_stream(cur)
                with torch.cuda.stream(streams[r]):
                    peers[r][r].copy_(t.view(world, n)[r], non_blocking=True)
            for r in range(1, world):
                cur.wait_stream(streams[r])
    report('symm push (rank 0, one stream each)', timed(push))
except Exception as e:                                                  # noqa: BLE001
    print('[rank %d] symmetric memory unavailable: %s: %s' % (rank, type(e).__name__, str(e)[:300]), flush=True)

dist.destroy_process_group()
