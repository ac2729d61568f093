"""Batch sharding of the model rollout over the GPUs of one box.

Every observation row is independent in every function of the path (the only reduction,
find_closest_point, runs over waypoints, not rows), so the path shards by contiguous row
blocks with NO data-path collective.  torch.distributed (NCCL over NVLink on GPUs, gloo in
the CPU tests) is used only to scatter a batch that lives on one rank and to gather the
per-row returns; path tables and configuration are replicated at construction.
"""
import torch
import torch.distributed as dist


def shard_bounds(B, world_size, rank):
    """Rows [lo, hi) of rank `rank`: contiguous blocks, the first B % W ranks one row longer."""
    base, rem = divmod(int(B), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(B, world_size):
    return [shard_bounds(B, world_size, r)[1] - shard_bounds(B, world_size, r)[0] for r in range(world_size)]


def _group_info(group=None):
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def scatter_rows(full, B, tail_shape, dtype, device, src=0, group=None):
    """Rank `src` holds `full` [B, *tail_shape]; every rank returns its row block.
    Ragged shards are padded to the longest one for the collective and trimmed after."""
    W, rank = _group_info(group)
    lo, hi = shard_bounds(B, W, rank)
    if W == 1:
        return full[lo:hi].to(device=device, dtype=dtype)
    longest = max(shard_sizes(B, W))
    recv = torch.empty((longest,) + tuple(tail_shape), dtype=dtype, device=device)
    chunks = None
    if rank == src:
        full = full.to(device=device, dtype=dtype)
        chunks = []
        for r in range(W):
            a, b = shard_bounds(B, W, r)
            c = torch.zeros((longest,) + tuple(tail_shape), dtype=dtype, device=device)
            c[:b - a] = full[a:b]
            chunks.append(c)
    dist.scatter(recv, chunks, src=src, group=group)
    return recv[:hi - lo]


def gather_rows(local, B, dst=0, group=None):
    """Inverse of scatter_rows for per-row results [b_local, ...]: rank `dst` returns [B, ...],
    the others None."""
    W, rank = _group_info(group)
    if W == 1:
        return local
    longest = max(shard_sizes(B, W))
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(W)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][:n] for r, n in enumerate(shard_sizes(B, W))], 0)


class ShardedRollout(object):
    """Data-parallel H-step rollout: rank `src` supplies [B, D] observations, [B] path indexes and
    an [H, B, 2] action tape; every rank rolls out its row block with `make_runner(b_local)`
    (a RolloutGraph-like object with load/run/out5) and rank `src` receives the per-row
    returns sum_t out5[t] as [B, 5]."""

    def __init__(self, make_runner, B, D, H, device, group=None):
        self.B, self.D, self.H, self.device, self.group = int(B), int(D), int(H), device, group
        W, rank = _group_info(group)
        lo, hi = shard_bounds(B, W, rank)
        self.lo, self.hi = lo, hi
        self.runner = make_runner(hi - lo)

    def scatter(self, obses=None, ref_indexes=None, tape=None, src=0):
        B = self.B
        obs = scatter_rows(obses, B, (self.D,), torch.float32, self.device, src, self.group)
        ref = scatter_rows(ref_indexes, B, (), torch.int32, self.device, src, self.group)
        tp = tape
        if tape is not None and _group_info(self.group)[1] == src:
            tp = tape.permute(1, 0, 2)                    # rows first for the row scatter
        tp = scatter_rows(tp, B, (self.H, 2), torch.float32, self.device, src, self.group)
        self.runner.load(obs, ref, tp.permute(1, 0, 2))

    def run(self):
        self.runner.run()

    def gather_returns(self, dst=0):
        ret = self.runner.out5.sum(0).t().contiguous()    # [b_local, 5]
        return gather_rows(ret, self.B, dst, self.group)
