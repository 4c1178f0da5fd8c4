"""Profiling target for the strip kernels: one PAMR iteration at several shapes (run under
ncu --metrics gpu__time_duration.sum with PAMR_B200_FORCE_STRIPS=3)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
SHAPES = [(321, 321, 16), (324, 324, 16), (328, 328, 16), (328, 328, 32), (321, 321, 4), (325, 325, 1), (648, 648, 4), (1028, 1028, 2)]
pamr = wseg_b200.PAMR(1, D6).to(dev)
for (H, W, B) in SHAPES:
    image = torch.rand((B, 3, H, W), device=dev); mask = torch.rand((B, 21, H, W), device=dev)
    for _ in range(2):
        out = pamr(image, mask)
    torch.cuda.synchronize()
    print(H, W, B, float(out[0, 0, 0, 0]))
