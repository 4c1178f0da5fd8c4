"""Class-balanced cross-entropy on the pseudo-labels, backed by libpamr_b200.so.

Mirrors  balanced_mask_loss_ce(mask, pseudo_gt, gt_labels, ignore_index=255)  of the reference
(models/SoftMaxAE.py:52-88, called at :258 / :311): `mask` are the decoder logits [B,C,h,w], `pseudo_gt`
the float one-hot-or-empty tensor [B,C,H,W], `gt_labels` the image-level labels [B,C-1]; returns the
per-sample loss [B] with autograd support for `mask` (the reference detaches everything else).
`balanced_mask_loss_ce_from_labels` is the fused entry for callers that already hold the uint8 label
map and per-class pixel counts of `stage.pseudo_labels(..., return_counts=True)` -- no float one-hot
tensor is ever built.  The bilinear up-sampling of the logits (:58) is evaluated inside the kernels.
"""
import torch

from . import _lib
from .pamr import _check_cuda_f32, _dev, _stream

IGNORE_INDEX = 255


class _MaskCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, label, counts, gt_labels):
        B, C, h, w = logits.shape
        H, W = int(label.shape[-2]), int(label.shape[-1])
        L = _lib.lib()
        need = L.pamr_mask_ce_workspace_bytes(B, C, h, w, H, W)
        ws = torch.empty((need,), dtype=torch.uint8, device=logits.device)
        loss = torch.empty((B,), dtype=torch.float32, device=logits.device)
        _lib.check(L.pamr_mask_ce_forward_f32(logits.data_ptr(), label.data_ptr(), counts.data_ptr(), gt_labels.data_ptr(),
                                              loss.data_ptr(), ws.data_ptr(), need, B, C, h, w, H, W, _dev(logits),
                                              _stream(logits.device)))
        ctx.save_for_backward(logits, label, ws)
        ctx.dims = (B, C, h, w, H, W, need)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        logits, label, ws = ctx.saved_tensors
        B, C, h, w, H, W, need = ctx.dims
        g = grad_loss.detach().to(dtype=torch.float32).contiguous()
        grad = torch.empty_like(logits)
        _lib.check(_lib.lib().pamr_mask_ce_backward_f32(logits.data_ptr(), label.data_ptr(), g.data_ptr(), grad.data_ptr(),
                                                        ws.data_ptr(), need, B, C, h, w, H, W, _dev(logits),
                                                        _stream(logits.device)))
        return grad, None, None, None


def _prep(mask, gt_labels):
    if not (isinstance(mask, torch.Tensor) and mask.is_cuda and mask.dtype == torch.float32 and mask.dim() == 4):
        raise RuntimeError("mask must be a CUDA float32 tensor [B,C,h,w] (no CPU fallback)")
    B, C = int(mask.shape[0]), int(mask.shape[1])
    gl = gt_labels.detach().to(device=mask.device, dtype=torch.float32).contiguous()
    if tuple(gl.shape) != (B, C - 1):
        raise RuntimeError("gt_labels must have shape [B, C-1] = %s, got %s" % ((B, C - 1), tuple(gl.shape)))
    return mask.contiguous(), gl


def labels_from_onehot(pseudo_gt):
    """argmax / ignore-255 (SoftMaxAE.py:61-66) and per-class pixel counts (:71-72) of a float
    one-hot-or-empty pseudo_gt [B,C,H,W]: returns (uint8 labels [B,H,W], int32 counts [B,C])."""
    pg = _check_cuda_f32("pseudo_gt", pseudo_gt)
    B, C, H, W = pg.shape
    label = torch.empty((B, H, W), dtype=torch.uint8, device=pg.device)
    counts = torch.empty((B, C), dtype=torch.int32, device=pg.device)
    _lib.check(_lib.lib().pamr_labels_from_onehot_f32(pg.data_ptr(), label.data_ptr(), counts.data_ptr(), B, C, H, W,
                                                      _dev(pg), _stream(pg.device)))
    return label, counts


def balanced_mask_loss_ce_from_labels(mask, label, counts, gt_labels):
    """Fused form: label uint8 [B,H,W] (255 = ignore), counts int32 [B,C] (pixels per class)."""
    mask, gl = _prep(mask, gt_labels)
    B, C = int(mask.shape[0]), int(mask.shape[1])
    if not (label.is_cuda and label.dtype == torch.uint8 and label.dim() == 3 and int(label.shape[0]) == B):
        raise RuntimeError("label must be a CUDA uint8 tensor [B,H,W]")
    if not (counts.is_cuda and counts.dtype == torch.int32 and tuple(counts.shape) == (B, C)):
        raise RuntimeError("counts must be a CUDA int32 tensor [B,C]")
    return _MaskCE.apply(mask, label.contiguous(), counts.contiguous(), gl)


def balanced_mask_loss_ce(mask, pseudo_gt, gt_labels, ignore_index=IGNORE_INDEX):
    """Reference signature (SoftMaxAE.py:52)."""
    if ignore_index != IGNORE_INDEX:
        raise RuntimeError("ignore_index is fixed at 255 (the uint8 label map's ignore value)")
    label, counts = labels_from_onehot(pseudo_gt.detach())
    return balanced_mask_loss_ce_from_labels(mask, label, counts, gt_labels)
