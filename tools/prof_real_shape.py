"""ncu target: three refine_and_label calls at stage_net's real call shape (image 321x321, masks 81x81, B=16)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
dev = "cuda:0"
B, C, H, W = 16, 21, 321, 321
h = int(sys.argv[1]) if len(sys.argv) > 1 else 81
pamr = wseg_b200.PAMR(10, [1, 2, 4, 8, 12, 24]).to(dev)
image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, h, h), device=dev), 1)
labels = (torch.rand((B, C - 1), device=dev) < 0.3).float(); labels[:, 0] = 1
for _ in range(3):
    wseg_b200.refine_and_label(pamr, image, mask, labels)
torch.cuda.synchronize()
