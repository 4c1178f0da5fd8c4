"""Batch-shard data parallelism for the PAMR hot path.

The reference scales with nn.DataParallel's dim-0 scatter/gather (train.py:112).  Here it is one
process per GPU (torchrun): every op on the path is independent per sample (pamr.py:46 views the
batch as B*K planes; the epilogue max is per (b,c), SoftMaxAE.py:31-35), so ranks take contiguous
batch slices with no halo exchange and the only collective is one all-gather of the compact
uint8 label maps (NCCL on GPUs; gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(batch, rank, world):
    """Contiguous slice [lo, hi) of a batch for `rank` of `world`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t, rank, world):
    lo, hi = shard_range(t.shape[0], rank, world)
    return t[lo:hi]


def gather_labels(local, batch, group=None):
    """All-gather per-rank label maps [b_r,H,W] (uint8) into the full [batch,H,W] on every rank.
    Shards of unequal size are padded to the largest one for the collective and trimmed after."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(batch, rank, world)
    if local.shape[0] != hi - lo:
        raise RuntimeError("rank %d holds %d samples, expected %d" % (rank, local.shape[0], hi - lo))
    if world == 1:
        return local
    cap = -(-batch // world)
    send = local
    if local.shape[0] != cap:
        send = local.new_full((cap,) + tuple(local.shape[1:]), 255)
        send[: local.shape[0]] = local
    out = local.new_empty((world * cap,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    if batch == world * cap:
        return out
    parts = []
    for r in range(world):
        a, b = shard_range(batch, r, world)
        parts.append(out[r * cap: r * cap + (b - a)])
    return torch.cat(parts, 0)


class ShardedPseudoLabeler:
    """Runs refine_and_label on this rank's batch slice and gathers the label maps."""

    def __init__(self, pamr, group=None):
        self.pamr = pamr
        self.group = group

    def __call__(self, image_raw, masks, labels, out_size=None, gather=True):
        from .stage import refine_and_label
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        B = image_raw.shape[0]
        lo, hi = shard_range(B, rank, world)
        if hi > lo:
            local = refine_and_label(self.pamr, image_raw[lo:hi], masks[lo:hi], labels[lo:hi], out_size)
        else:  # B < world: this rank's shard is empty; it still takes part in the gather (the C API rejects B < 1)
            H, W = (int(out_size[0]), int(out_size[1])) if out_size is not None else tuple(image_raw.shape[-2:])
            local = torch.empty((0, H, W), dtype=torch.uint8, device=image_raw.device)
        if not gather or world == 1:
            return local
        return gather_labels(local, B, self.group)
