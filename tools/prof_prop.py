"""Profiling target: a few single propagation iterations at config-2 shape (run under ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
B, C, H, W = int(os.environ.get("PROF_B", 16)), 21, int(os.environ.get("PROF_H", 321)), int(os.environ.get("PROF_W", 321))
image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, H, W), device=dev), 1)
aff = wseg_b200.local_affinity(image, D6)
for _ in range(int(os.environ.get("PROF_N", 4))):
    out = wseg_b200.propagate(aff, mask, D6, 2)
torch.cuda.synchronize()
print("ok", float(out.sum()))
