"""Time balanced_mask_loss_ce forward + backward at the training shapes and, for
context, a plain torch transcription of the same math on the same GPU (cuDNN/ATen kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F, wseg_b200
dev = "cuda:0"

def torch_loss(mask, pseudo_gt, gt_labels):  # the reference's formulas on ATen kernels
    mask = F.interpolate(mask, size=pseudo_gt.shape[-2:], mode="bilinear", align_corners=True)
    mask_gt = torch.argmax(pseudo_gt, 1)
    mask_gt[pseudo_gt.sum(1) < 1.] = 255
    bs, c, h, w = pseudo_gt.shape
    n = pseudo_gt.view(bs, c, -1).sum(-1)
    tot = n.sum(-1, keepdim=True)
    cw = (tot - n) / (1 + tot)
    cw = (pseudo_gt * cw[:, :, None, None]).sum(1).view(bs, -1)
    loss = F.cross_entropy(mask, mask_gt, ignore_index=255, reduction="none").view(bs, -1)
    bw = ((gt_labels.sum(-1) + 1) == (n > 0).float().sum(-1)).float()
    return bw * (cw * loss).mean(-1)

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B, C, H, W = 16, 21, 321, 321
for (h, w) in [(41, 41), (81, 81), (321, 321)]:
    g = torch.Generator(device=dev).manual_seed(0)
    logits = (2 * torch.randn((B, C, h, w), generator=g, device=dev)).requires_grad_(True)
    masks = torch.softmax(2 * torch.randn((B, C, H, W), generator=g, device=dev), 1)
    lab, cnt, onehot = None, None, None
    lab, onehot, cnt = wseg_b200.pseudo_labels(masks, None, None, None, return_onehot=True, return_counts=True)
    gl = (cnt[:, 1:] > 0).float()
    def ours():
        loss = wseg_b200.balanced_mask_loss_ce_from_labels(logits, lab, cnt, gl)
        loss.sum().backward(); logits.grad = None
    def ours_ref_sig():
        loss = wseg_b200.balanced_mask_loss_ce(logits, onehot, gl)
        loss.sum().backward(); logits.grad = None
    def aten():
        loss = torch_loss(logits, onehot, gl)
        loss.sum().backward(); logits.grad = None
    a = wseg_b200.balanced_mask_loss_ce_from_labels(logits, lab, cnt, gl); b = torch_loss(logits, onehot, gl)
    print("logits %3dx%-3d -> labels %dx%d  B=%d C=%d: fused labels path %.3f ms, reference signature (float one-hot in) %.3f ms, "
          "ATen transcription %.3f ms   (loss rel diff %.1e)" % (h, w, H, W, B, C, t(ours), t(ours_ref_sig), t(aten),
          float(((a - b).abs() / b.abs().clamp_min(1e-12)).max())))
