"""Pinned host -> device copy bandwidth for the e2e inputs of config 2, on N GPUs at once (is bench.py's e2e bounded by
the host side?).  Run as  torchrun --nproc-per-node N tools/h2d_bw.py  (or plain python for one GPU): every rank copies
its own 158 MB (image + full-resolution mask) per step to its own GPU; rank 0 prints per-rank and aggregate GB/s, with
and without binding each rank to the CPUs NVML reports next to its GPU."""
import os, sys
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
bind = os.environ.get("H2D_BIND", "1") == "1"
aff = "unbound"
if bind:
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        aff = "cpus %s" % (sorted(os.sched_getaffinity(0))[:1] + ["..."] + sorted(os.sched_getaffinity(0))[-1:])
    except Exception as e:
        aff = "bind failed: %r" % (e,)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
B, K, C, H, W = 16, 3, 21, 321, 321
h_img = torch.rand((B, K, H, W)).pin_memory(); h_msk = torch.rand((B, C, H, W)).pin_memory()
d_img = torch.empty_like(h_img, device=dev); d_msk = torch.empty_like(h_msk, device=dev)
def t(fn, n=30):
    for _ in range(3): fn()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def both():
    d_img.copy_(h_img, non_blocking=True); d_msk.copy_(h_msk, non_blocking=True)
ms = t(both)
nbytes = h_img.numel() * 4 + h_msk.numel() * 4
gbs = nbytes / ms / 1e6
if world > 1:
    all_ = [None] * world
    dist.all_gather_object(all_, (rank, gbs, aff))
else:
    all_ = [(0, gbs, aff)]
if rank == 0:
    print("N=%d concurrent H2D of %.1f MB per rank (%s): per rank %s GB/s; aggregate %.1f GB/s; slowest rank %.3f ms per step" % (
        world, nbytes / 1e6, "bound" if bind else "unbound", ["%.1f" % g for _, g, _ in all_], sum(g for _, g, _ in all_),
        nbytes / min(g for _, g, _ in all_) / 1e6), flush=True)
    print("   affinity:", [a for _, _, a in all_][:2], "...", flush=True)
if world > 1:
    dist.destroy_process_group()
