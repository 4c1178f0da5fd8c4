"""Profiling target: the forward path's affinity kernels at config-2 shape (PAMR with one iteration)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
B, C = int(os.environ.get("PROF_B", 16)), 21
H, W = int(os.environ.get("PROF_H", 321)), int(os.environ.get("PROF_W", 321))
image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, H, W), device=dev), 1)
pamr = wseg_b200.PAMR(1, D6).to(dev)
for _ in range(3):
    out = pamr(image, mask)
torch.cuda.synchronize()
print("ok", float(out.sum()))
