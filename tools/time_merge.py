"""Time the multi-scale merge + prediction at the shipped inference shape: 8 scale/flip
views of a 21-class score map padded to 1024x1024 (configs/voc_resnet38.yaml), one 500x375 image."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
dev = "cuda:0"
S, C, Hp, Wp, H, W = 8, 21, 1024, 1024, 375, 500
scales = [1, 0.5, 1.5, 2.0]
masks = torch.zeros((S, C, Hp, Wp), device=dev)
pads = []
for s in range(S):
    h, w = int(round(H * scales[s // 2])), int(round(W * scales[s // 2]))
    pt, pl = (Hp - h) // 2, (Wp - w) // 2
    masks[s, :, pt:pt + h, pl:pl + w] = torch.softmax(2 * torch.randn((C, h, w), device=dev), 0)
    pads.append((pt, pl, h, w))
labels = (torch.rand((C - 1,), device=dev) < 0.3).float()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
h_pred = torch.empty((H, W), dtype=torch.uint8).pin_memory()
def ours():
    h_pred.copy_(wseg_b200.merge_and_predict(masks, labels, pads, (H, W), 0.3, flip=True, bg_pow=3), non_blocking=True)
h_masks = torch.empty(masks.shape, dtype=torch.float32).pin_memory()
def ref_transfer():  # what the reference does first (infer_val.py:124): the scores go to the host
    h_masks.copy_(masks, non_blocking=True)
print("merge + predict on the device, uint8 map to pinned host: %.3f ms;  D2H of the [8,21,1024,1024] scores alone "
      "(the reference's first step before its numpy merge): %.3f ms" % (t(ours), t(ref_transfer, 5)))
