"""Profiling target: a few PAMR forwards (+ epilogue) at config-2 shape (run under ncu for a launch list)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
B, C = int(os.environ.get("PROF_B", 16)), 21
H, W = int(os.environ.get("PROF_H", 321)), int(os.environ.get("PROF_W", 321))
h, w = int(os.environ.get("PROF_h", H)), int(os.environ.get("PROF_w", W))
image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, h, w), device=dev), 1)
labels = (torch.rand((B, C - 1), device=dev) < 0.3).float(); labels[:, 0] = 1
pamr = wseg_b200.PAMR(10, D6).to(dev)
for _ in range(int(os.environ.get("PROF_N", 3))):
    out = wseg_b200.refine_and_label(pamr, image, mask, labels)
torch.cuda.synchronize()
print("ok", int(out.sum()))
