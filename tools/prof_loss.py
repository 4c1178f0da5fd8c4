"""ncu target: balanced_mask_loss_ce forward + backward on the label maps at the training shapes (logits 81x81 and
41x41, labels 321x321, B=16), two calls each."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
dev = "cuda:0"
B, C, H, W = 16, 21, 321, 321
g = torch.Generator(device=dev).manual_seed(0)
masks = torch.softmax(2 * torch.randn((B, C, H, W), generator=g, device=dev), 1)
lab, cnt = wseg_b200.pseudo_labels(masks, None, None, None, return_counts=True)
gl = (cnt[:, 1:] > 0).float()
for h in (81, 41):
    logits = (2 * torch.randn((B, C, h, h), generator=g, device=dev)).requires_grad_(True)
    for _ in range(2):
        loss = wseg_b200.balanced_mask_loss_ce_from_labels(logits, lab, cnt, gl)
        loss.sum().backward(); logits.grad = None
torch.cuda.synchronize()
