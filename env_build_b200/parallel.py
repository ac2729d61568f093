"""Batch sharding of the model rollout over the GPUs of one box.

Every observation row is independent in every function of the path (the only reduction,
find_closest_point, runs over waypoints, not rows), so the path shards by contiguous row
blocks with NO data-path collective.  torch.distributed (NCCL over NVLink on GPUs, gloo in
the CPU tests) is used only to scatter a batch that lives on one rank and to gather the
per-row returns; path tables and configuration are replicated at construction.

Two ways to feed a `ShardedRollout`:

* `scatter(obses, ref_indexes, tape)`: any row count, inputs in the API layout ([B, D], [B],
  [H, B, 2]).  Ragged shards are padded to the longest one for the collective.
* `stage(...)` once on the source rank, then `scatter_staged(...)` per rollout: the source keeps the
  batch in the layout the ranks' static buffers have (padded observation rows, action tape blocked
  by rank), so the collectives send VIEWS of it straight into those buffers: no pad, no copy, no
  permute on either side.  Needs B % world_size == 0.  With `slots=2` the next rollout's inputs can
  travel (on a side stream) while the current one computes.

The staged exchange has two transports.  `exchange='nccl'`: one `dist.scatter` (send / recv kernels, which
need SMs the persistent step kernel fills).  `exchange='peer'`: the staged batch lives in a buffer that every
rank of the box maps (torch symmetric memory: CUDA VMM allocations exported over NVLink peer mappings), and
every rank PULLS its block into its inbox with the copy engines -- no kernel on any SM but a one-block
barrier, and the source's NVLink egress runs at the peer-copy rate instead of NCCL's send rate.
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
    Equal shards are sent as views of `full`; ragged shards are padded to the longest one for the
    collective and trimmed after."""
    W, rank = _group_info(group)
    lo, hi = shard_bounds(B, W, rank)
    if W == 1:
        return full[lo:hi].to(device=device, dtype=dtype)
    sizes = shard_sizes(B, W)
    longest = max(sizes)
    recv = torch.empty((longest,) + tuple(tail_shape), dtype=dtype, device=device)
    chunks = None
    if rank == src:
        full = full.to(device=device, dtype=dtype).contiguous()
        if min(sizes) == longest:
            chunks = list(full.split(longest, 0))
        else:
            chunks = []
            for r in range(W):
                a, b = shard_bounds(B, W, r)
                if b - a == longest:
                    chunks.append(full[a:b])
                else:
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
    sizes = shard_sizes(B, W)
    longest = max(sizes)
    if local.shape[0] != longest:
        pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
        local = pad
    out = torch.empty((W, longest) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) \
        if rank == dst else None
    dist.gather(local.contiguous(), list(out.unbind(0)) if rank == dst else None, dst=dst, group=group)
    if rank != dst:
        return None
    if min(sizes) == longest:
        return out.reshape((B,) + tuple(local.shape[1:]))
    return torch.cat([out[r, :n] for r, n in enumerate(sizes)], 0)


class StagedBatch(object):
    """A global batch on the source rank in scatter-ready layout (see ShardedRollout.stage)."""

    def __init__(self, obs_store, ref, tape, b, ld):
        self.obs_store, self.ref, self.tape, self.b, self.ld = obs_store, ref, tape, b, ld


class ShardedRollout(object):
    """Data-parallel H-step rollout: rank `src` supplies [B, D] observations, [B] path indexes and
    an [H, B, 2] action tape; every rank rolls out its row block with `make_runner(b_local)`
    (a RolloutGraph-like object with load/run/out5) and rank `src` receives the per-row
    returns sum_t out5[t] as [B, 5].  `slots` runners per rank allow one rollout's exchange to overlap
    another's compute."""

    def __init__(self, make_runner, B, D, H, device, group=None, slots=1, exchange='nccl', src=0):
        """exchange='peer' (CUDA, one box, B % world_size == 0; collective: every rank constructs it): the
        staged batch of rank `src` is pulled over peer mappings, see the module docstring.  Every rank maps a
        window of the staged batch's size (only `src`'s is read)."""
        self.B, self.D, self.H, self.device, self.group = int(B), int(D), int(H), device, group
        W, rank = _group_info(group)
        self.world, self.rank = W, rank
        lo, hi = shard_bounds(B, W, rank)
        self.lo, self.hi = lo, hi
        self.runners = [make_runner(hi - lo) for _ in range(int(slots))]
        self.runner = self.runners[0]
        self._gather_out = {}
        if exchange not in ('nccl', 'peer'):
            raise ValueError("exchange is 'nccl' or 'peer'")
        self.exchange, self.src = exchange, int(src)
        self._window = self._peer = self._src_blocks = None
        if exchange == 'peer' and W > 1:
            if B % W:
                raise ValueError('the staged scatter needs B %% world_size == 0 (B=%d, W=%d)' % (B, W))
            import torch.distributed._symmetric_memory as symm
            n_in = self.runner.inbox.numel()
            self._window = symm.empty(W * n_in, dtype=torch.float32, device=device)
            name = (group if group is not None else dist.group.WORLD).group_name
            self._peer = symm.rendezvous(self._window, name)
            self._src_blocks = self._peer.get_buffer(self.src, (W, n_in), torch.float32)

    # -- generic path -------------------------------------------------------------------------
    def scatter(self, obses=None, ref_indexes=None, tape=None, src=0, slot=0, has_ref=True, has_tape=True):
        """`has_ref` / `has_tape` must agree on every rank (mode != 'training' needs no path indexes; a
        closed-loop runner needs no tape): the matching collective is skipped when False."""
        B = self.B
        obs = scatter_rows(obses, B, (self.D,), torch.float32, self.device, src, self.group)
        ref = scatter_rows(ref_indexes, B, (), torch.int32, self.device, src, self.group) if has_ref else None
        tp = None
        if has_tape:
            tp = tape
            if tape is not None and self.rank == src:
                tp = tape.permute(1, 0, 2)                # rows first for the row scatter
            tp = scatter_rows(tp, B, (self.H, 2), torch.float32, self.device, src, self.group).permute(1, 0, 2)
        self.runners[slot].load(obs, ref, tp)

    # -- staged path: views straight into the runners' static buffers -----------------------------
    def stage(self, obses, ref_indexes, tape):
        """On the source rank: the global batch as one buffer of W per-rank blocks, each laid out like a
        runner's `inbox` (padded observation rows | action tape [H, b, 2] | path indexes).  One-time layout
        work, outside the per-rollout exchange.  With exchange='peer' the blocks are written into the window the
        other ranks pull from: stage the next batch only after the returns of every rollout that pulled the
        previous one have been gathered."""
        import numpy as np
        from .dynamics_and_models import padded_rows

        def to_device(x, dtype=torch.float32):
            t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
            return t.to(device=self.device, dtype=dtype)
        B, W = self.B, self.world
        if B % W:
            raise ValueError('the staged scatter needs B %% world_size == 0 (B=%d, W=%d)' % (B, W))
        b = B // W
        r = self.runner
        n_in = r.inbox.numel()
        ld, off = r.obs0.stride(0), r.obs0.storage_offset()
        n_obs = n_in - self.H * b * 2 - b
        if self._window is not None:
            if self.rank != self.src:
                raise ValueError('stage() runs on the source rank (%d)' % self.src)
            out = self._window.view(W, n_in).zero_()
        else:
            out = torch.zeros((W, n_in), dtype=torch.float32, device=self.device)
        obs, ref, tp = to_device(obses), to_device(ref_indexes, torch.int32).reshape(B), to_device(tape)
        for k in range(W):
            blk = out[k]
            blk.as_strided((b, self.D), (ld, 1), blk.storage_offset() + off).copy_(obs[k * b:(k + 1) * b])
            blk[n_obs:n_obs + self.H * b * 2].view(self.H, b, 2).copy_(tp[:, k * b:(k + 1) * b])
            blk[n_obs + self.H * b * 2:].view(torch.int32).copy_(ref[k * b:(k + 1) * b])
        return StagedBatch(out, None, None, b, ld)

    def scatter_staged(self, staged=None, slot=0, src=0):
        """ONE collective: the send buffers are the rows of the staged batch, the receive buffer is the
        runner's own inbox (RolloutGraph.obs0 / .tape / .ref are views of it)."""
        r = self.runners[slot]
        if self.world == 1:
            r.inbox.copy_(staged.obs_store[0])
            return
        if self._peer is not None:
            # stream-ordered barrier over the peer mappings (one block): the source's staging writes precede it,
            # every rank's pull follows it.  The gather of the returns tells the source when the window may be
            # overwritten.  The pull itself is a device-to-device copy whose source is NVLink-mapped memory.
            if src != self.src:
                raise ValueError('the window belongs to rank %d' % self.src)
            self._peer.barrier(channel=slot)
            r.inbox.copy_(self._src_blocks[self.rank], non_blocking=True)
            return
        dist.scatter(r.inbox, list(staged.obs_store.unbind(0)) if self.rank == src else None, src=src, group=self.group)

    def run(self, slot=0):
        self.runners[slot].run()

    def gather_returns(self, dst=0, slot=0):
        ret = self.runners[slot].out5.sum(0).t().contiguous()    # [b_local, 5]
        return gather_rows(ret, self.B, dst, self.group)
