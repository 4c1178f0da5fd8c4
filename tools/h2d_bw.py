"""Pinned host -> device copy bandwidth for the e2e inputs of config 2 (is bench.py's e2e PCIe-bound?)."""
import torch
dev = "cuda:0"
B, K, C, H, W = 16, 3, 21, 321, 321
h_img = torch.rand((B, K, H, W)).pin_memory(); h_msk = torch.rand((B, C, H, W)).pin_memory()
d_img = torch.empty_like(h_img, device=dev); d_msk = torch.empty_like(h_msk, device=dev)
s = torch.cuda.Stream()
def t(fn, n=20):
    for _ in range(3): fn()
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
print("H2D %.1f MB in %.3f ms = %.1f GB/s" % (nbytes / 1e6, ms, nbytes / ms / 1e6))
big = torch.empty((nbytes // 4,), dtype=torch.float32).pin_memory(); dbig = torch.empty_like(big, device=dev)
ms = t(lambda: dbig.copy_(big, non_blocking=True))
print("single buffer: %.3f ms = %.1f GB/s" % (ms, nbytes / ms / 1e6))
