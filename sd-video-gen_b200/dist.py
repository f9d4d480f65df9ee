"""Clip-sharded multi-GPU rollout: one process per GPU, clips are independent, no
data-path collective - each rank rolls out a contiguous block of clips and only the predictions are gathered.

The one subtlety is the reference's positional encoding, which is indexed by the clip's position in the batch
(models/positional_encoding.py:35): the reference can only process <= 64 clips per call, so a large batch is
by definition processed in chunks of 64 and clip i sees PE[i mod 64].  Shards keep that by passing
``pe_index = global_clip_index mod 64``.
"""
import torch


def shard_bounds(n_clips, rank, world_size, multiple=64):
    """Contiguous [start, stop) of clips for `rank`; block sizes are multiples of `multiple` where possible."""
    blocks = (n_clips + multiple - 1) // multiple
    per, extra = divmod(blocks, world_size)
    b0 = rank * per + min(rank, extra)
    b1 = b0 + per + (1 if rank < extra else 0)
    return min(b0 * multiple, n_clips), min(b1 * multiple, n_clips)


def pe_index_for(start, stop, device=None):
    return (torch.arange(start, stop, dtype=torch.int64) % 64).to(dtype=torch.int32, device=device)


def rollout_sharded(rollout_fn, ctx, n_pred, window=5, *, rank=None, world_size=None, gather=True, **kw):
    """Run ``rollout_fn(ctx_shard, n_pred, window, pe_index=..., **kw)`` on this rank's clips and gather.

    ctx is the FULL (B,C,E) batch (same on every rank - synthetic data is generated from one seed) or already
    this rank's shard when ``gather=False``.  Returns (B,n_pred,E) on every rank when gathering."""
    import torch.distributed as dist
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if world_size > 1 else 0
    B = ctx.size(0)
    s, e = shard_bounds(B, rank, world_size)
    local = rollout_fn(ctx[s:e], n_pred, window, pe_index=pe_index_for(s, e, ctx.device), **kw) if e > s else \
        ctx.new_zeros((0, n_pred, ctx.size(2)))
    if not gather or world_size == 1:
        return local
    sizes = [shard_bounds(B, r, world_size) for r in range(world_size)]
    pad = max(b - a for a, b in sizes)
    buf = local.new_zeros((pad, n_pred, local.size(2)))
    buf[: e - s] = local
    outs = [torch.empty_like(buf) for _ in range(world_size)]
    dist.all_gather(outs, buf)
    return torch.cat([o[: b - a] for o, (a, b) in zip(outs, sizes)], 0)
