"""Time one propagation iteration for a few shapes (tuned vs generic via PAMR_B200_FORCE_GENERIC)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
shapes = [(16, 21, 321, 321), (16, 21, 320, 320), (2, 21, 1024, 1024), (16, 21, 81, 81)]
for (B, C, H, W) in shapes:
    image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, H, W), device=dev), 1)
    aff = wseg_b200.local_affinity(image, D6)
    for iters in (1, 10):
        for _ in range(2): wseg_b200.propagate(aff, mask, D6, iters)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = wseg_b200._lib.launch_count()
        e0.record()
        for _ in range(5): wseg_b200.propagate(aff, mask, D6, iters)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gbs = 360.0 * B * H * W * iters / (ms * 1e-3) / 1e9
        print("%s iters=%2d: %8.3f ms  (%.3f ms/iter)  %7.1f GB/s algorithmic = %4.1f%% of 6543; launches/call %d"
              % ((B, C, H, W), iters, ms, ms / iters, gbs, gbs / 65.431, (wseg_b200._lib.launch_count() - n0) // 5), flush=True)
