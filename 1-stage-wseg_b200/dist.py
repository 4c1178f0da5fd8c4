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


class OverlappedLabelGather:
    """All-gather of the per-rank uint8 label maps on a SIDE stream, so that the collective of step i runs next to
    the kernels of step i+1 ("overlap with the final epilogue tile wave").

    submit(local) enqueues, on the side stream, an all_gather_into_tensor of `local` ([b,H,W] uint8, the same shape
    on every rank) into one of `depth` rotating output buffers -- ordered after everything enqueued so far on the
    caller's current stream -- and returns that buffer ([world*b,H,W]).  With host_out (pinned) the gathered maps are
    also copied to the host on the side stream.  The caller's stream does NOT wait: call wait() before reading the
    returned buffer (or host_out) and before the buffer comes round again (depth submits later).
    On CPU tensors (gloo, the unit tests) the gather is synchronous."""

    def __init__(self, device, group=None, depth=2):
        self.device = torch.device(device)
        self.group = group
        self.depth = int(depth)
        self.cuda = self.device.type == "cuda"
        self.side = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._bufs = {}
        self._n = 0

    def _buffer(self, local, world):
        key = (tuple(local.shape), local.dtype)
        if key not in self._bufs:
            if self.cuda:  # the old buffers may still be in flight
                self.side.synchronize()
            shape = (world * local.shape[0],) + tuple(local.shape[1:])
            self._bufs = {key: [torch.empty(shape, dtype=local.dtype, device=local.device) for _ in range(self.depth)]}
        buf = self._bufs[key][self._n % self.depth]
        self._n += 1
        return buf

    def submit(self, local, host_out=None):
        world = dist.get_world_size(self.group)
        out = self._buffer(local, world)
        if not self.cuda:
            dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
            if host_out is not None:
                host_out.copy_(self._host_view(out, local, host_out))
            return out
        main = torch.cuda.current_stream(self.device)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            dist.all_gather_into_tensor(out, local, group=self.group)
            if host_out is not None:
                host_out.copy_(self._host_view(out, local, host_out), non_blocking=True)
        local.record_stream(self.side)  # the allocator must not hand `local` out again before the gather has read it
        return out

    def _host_view(self, out, local, host_out):
        if host_out.shape[0] == out.shape[0]:
            return out
        b, rank = local.shape[0], dist.get_rank(self.group)
        if host_out.shape[0] != b:
            raise RuntimeError("host_out must hold all %d gathered maps or this rank's %d" % (out.shape[0], b))
        return out[rank * b: (rank + 1) * b]

    def wait(self):
        """Everything submitted so far becomes visible to the caller's current stream."""
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
