"""Pseudo-label epilogue of stage_net around PAMR, backed by libpamr_b200.so.

Mirrors the helpers every reference model file carries (paths relative to the reference root):
  run_pamr            models/SoftMaxAE.py:176-179   image -> mask size, then PAMR
  _rescale_and_clean  models/SoftMaxAE.py:263-268   bilinear to the image size, masks[:,1:] *= labels
  pseudo_gtmask       models/SoftMaxAE.py:29-50     per-class thresholds, ambiguity removal
  argmax / ignore     models/SoftMaxAE.py:61-67     label map with 255 = ignore
plus `pseudo_labels` / `refine_and_label`, the fused path that never materialises the
up-sampled float tensors (what bench.py times as "PAMR + clean/argmax epilogue").
"""
import torch

from . import _lib
from .pamr import PAMR, _check_cuda_f32, _dev, _stream, resize_bilinear

IGNORE_INDEX = 255


VOC_MEAN, VOC_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)  # datasets/pascal_voc.py:69-70


def denorm_resize(image_norm, mean=VOC_MEAN, std=VOC_STD, size=None):
    """`denorm(image.clone())` (datasets/pascal_voc.py:85-101, train.py:120) and the bilinear resize of
    run_pamr (SoftMaxAE.py:177) in one pass: returns a new tensor [B,K,H,W]; the input is not modified."""
    import ctypes
    x = _check_cuda_f32("image", image_norm)
    B, K, h, w = x.shape
    H, W = _size_of(size) if size is not None else (h, w)
    if len(mean) != K or len(std) != K:
        raise RuntimeError("mean / std must have one entry per image channel (%d)" % K)
    out = torch.empty((B, K, H, W), dtype=torch.float32, device=x.device)
    c_mean, c_std = (ctypes.c_float * K)(*[float(v) for v in mean]), (ctypes.c_float * K)(*[float(v) for v in std])
    _lib.check(_lib.lib().pamr_denorm_resize_f32(x.data_ptr(), ctypes.cast(c_mean, ctypes.c_void_p), ctypes.cast(c_std, ctypes.c_void_p),
                                                 out.data_ptr(), B, K, h, w, H, W, _dev(x), _stream(x.device)))
    return out


def run_pamr(pamr, im, mask, denorm=None):
    """run_pamr(self, im, mask) with self._aff == pamr (SoftMaxAE.py:176-179).  With denorm=(mean, std) `im`
    is the NORMALISED network input and train.py:120's denorm is folded into the resize."""
    if denorm is not None:
        im = denorm_resize(im, denorm[0], denorm[1], mask.shape[-2:])
    elif tuple(im.shape[-2:]) != tuple(mask.shape[-2:]):
        im = resize_bilinear(im, mask.shape[-2:])
    return pamr(im, mask)


def _labels_arg(labels, B, C, device):
    if labels is None:
        return None
    labels = labels.detach().to(device=device, dtype=torch.float32).contiguous()
    if tuple(labels.shape) != (B, C - 1):
        raise RuntimeError("labels must have shape [B, C-1] = %s, got %s" % ((B, C - 1), tuple(labels.shape)))
    return labels


def _size_of(image_or_size):
    if isinstance(image_or_size, torch.Tensor):
        return int(image_or_size.shape[-2]), int(image_or_size.shape[-1])
    return int(image_or_size[0]), int(image_or_size[1])


def rescale_and_clean(masks, image, labels, return_class_max=False):
    """_rescale_and_clean(self, masks, image, labels) (SoftMaxAE.py:263-268): returns a NEW tensor
    [B,C,H,W]; `masks` is not modified.  `image` may be a tensor or an (H, W) pair."""
    masks = _check_cuda_f32("masks", masks)
    B, C, h, w = masks.shape
    H, W = _size_of(image)
    lab = _labels_arg(labels, B, C, masks.device)
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=masks.device)
    cmax = torch.empty((B, C), dtype=torch.int32, device=masks.device) if return_class_max else None
    _lib.check(_lib.lib().pamr_clean_f32(masks.data_ptr(), lab.data_ptr() if lab is not None else None, out.data_ptr(),
                                         cmax.data_ptr() if cmax is not None else None, B, C, h, w, H, W, _dev(masks),
                                         _stream(masks.device)))
    return (out, cmax) if return_class_max else out


def pseudo_labels(masks, labels=None, size=None, class_max=None, cutoff_top=0.6, cutoff_low=0.2, cutoff_bg=0.7,
                  return_onehot=False, return_counts=False):
    """Fused _rescale_and_clean -> pseudo_gtmask -> argmax/ignore (SoftMaxAE.py:263-268, 29-50, 61-67).

    masks [B,C,h,w] (PAMR output, un-gated), labels [B,C-1] or None, size = target (H, W) or None.
    class_max: the un-gated per-class max returned by PAMR.forward(..., return_class_max=True)
    (only usable without a resize); otherwise it is computed here.
    Returns uint8 labels [B,H,W] (255 = ignore) and, on request, the float one-hot pseudo_gt
    [B,C,H,W] and int32 per-class pixel counts [B,C]."""
    masks = _check_cuda_f32("masks", masks)
    B, C, h, w = masks.shape
    H, W = _size_of(size) if size is not None else (h, w)
    lab = _labels_arg(labels, B, C, masks.device)
    L = _lib.lib()
    dev, st = _dev(masks), _stream(masks.device)
    resize = (h, w) != (H, W)
    gated = 1
    if class_max is None or resize:
        class_max = torch.empty((B, C), dtype=torch.int32, device=masks.device)
        _lib.check(L.pamr_clean_f32(masks.data_ptr(), lab.data_ptr() if lab is not None else None, None,
                                    class_max.data_ptr(), B, C, h, w, H, W, dev, st))
    else:
        gated = 0
    out = torch.empty((B, H, W), dtype=torch.uint8, device=masks.device)
    onehot = torch.empty((B, C, H, W), dtype=torch.float32, device=masks.device) if return_onehot else None
    counts = torch.empty((B, C), dtype=torch.int32, device=masks.device) if return_counts else None
    _lib.check(L.pamr_pseudo_labels_f32(
        masks.data_ptr(), lab.data_ptr() if lab is not None else None, class_max.data_ptr(), out.data_ptr(),
        onehot.data_ptr() if onehot is not None else None, counts.data_ptr() if counts is not None else None, B, C, h, w,
        H, W, float(cutoff_bg), float(cutoff_top), float(cutoff_low), gated, dev, st))
    res = (out,)
    if return_onehot:
        res += (onehot,)
    if return_counts:
        res += (counts,)
    return res if len(res) > 1 else out


def pseudo_gtmask(mask, cutoff_top=0.6, cutoff_low=0.2, eps=1e-8):
    """pseudo_gtmask(mask, cutoff_top, cutoff_low, eps) (SoftMaxAE.py:29-50): float one-hot-or-empty
    [B,C,H,W].  (`eps` is unused in the reference as well.)"""
    _, onehot = pseudo_labels(mask, None, None, None, cutoff_top, cutoff_low, 0.7, return_onehot=True)
    return onehot


def labels_from_pseudo_gt(pseudo_gt, ignore_index=IGNORE_INDEX):
    """argmax + ignore (SoftMaxAE.py:61-67) for callers that hold a float pseudo_gt already: int64 label
    map like the reference's (one kernel, pamr_labels_from_onehot_f32; the uint8 map widened on the device)."""
    pg = _check_cuda_f32("pseudo_gt", pseudo_gt)
    B, C, H, W = pg.shape
    lab = torch.empty((B, H, W), dtype=torch.uint8, device=pg.device)
    _lib.check(_lib.lib().pamr_labels_from_onehot_f32(pg.data_ptr(), lab.data_ptr(), None, B, C, H, W, _dev(pg),
                                                      _stream(pg.device)))
    out = lab.long()
    if ignore_index != IGNORE_INDEX:
        out[lab == IGNORE_INDEX] = ignore_index
    return out


def refine_and_label(pamr, image_raw, masks, labels, out_size=None, return_masks=False, return_counts=False, denorm=None):
    """Sequence A of stage_net (SoftMaxAE.py:250-259) in as few passes as possible:
    run_pamr(image_raw, masks) -> _rescale_and_clean(., labels) -> pseudo_gtmask -> label map.

    image_raw [B,K,Hi,Wi], masks [B,C,h,w] (softmax scores), labels [B,C-1].  out_size defaults
    to the image size.  Returns uint8 labels [B,H,W]; with return_masks also the refined
    (un-gated, mask-resolution) masks_dec."""
    if not isinstance(pamr, PAMR):
        raise TypeError("pamr must be a wseg_b200.PAMR module")
    H, W = _size_of(out_size) if out_size is not None else _size_of(image_raw)
    h, w = int(masks.shape[-2]), int(masks.shape[-1])
    if denorm is not None:  # image_raw is the normalised input: fold train.py:120's denorm into the resize
        im = denorm_resize(image_raw, denorm[0], denorm[1], (h, w))
    else:
        im = resize_bilinear(image_raw, (h, w)) if tuple(image_raw.shape[-2:]) != (h, w) else image_raw
    if (h, w) == (H, W):
        dec, cmax = pamr(im, masks, return_class_max=True)  # class max fused into the last iteration
        res = pseudo_labels(dec, labels, None, cmax, return_counts=return_counts)
    else:
        dec = pamr(im, masks)
        res = pseudo_labels(dec, labels, (H, W), None, return_counts=return_counts)
    if not return_masks:
        return res
    return (res + (dec,)) if isinstance(res, tuple) else (res, dec)


class HostPipeline:
    """refine_and_label for HOST (pinned) inputs, as a data-loader-side caller uses it: inputs are
    copied host->device on a dedicated copy stream into one of two staging buffer sets, the kernels
    run on the caller's current stream, and the uint8 label maps are copied back to (pinned) host
    memory.  The two buffer sets alternate across chunks AND across calls, so the copy of the next
    batch (or chunk) overlaps the kernels of the current one -- the pattern of a training loop that
    prefetches its next batch.  The call is asynchronous: synchronise the current stream (or the
    device) before reading the returned host tensor.  This is bench.py's `e2e` path.

    chunks > 1 additionally splits one batch so that its own copy overlaps its own compute (useful
    for a single large batch; for B=16 at 321x321 the smaller launches cost more than they hide)."""

    def __init__(self, pamr, device, chunks=1):
        if not isinstance(pamr, PAMR):
            raise TypeError("pamr must be a wseg_b200.PAMR module")
        self.pamr, self.device, self.chunks = pamr, torch.device(device), int(chunks)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._bufs = None
        self._key = None
        self._n = 0  # staging buffer sets handed out so far (parity selects the set)
        self._ready = [torch.cuda.Event() for _ in range(2)]  # set j: inputs copied in
        self._freed = [torch.cuda.Event() for _ in range(2)]  # set j: consumed by the kernels

    def _buffers(self, image, masks, labels, n):
        key = (tuple(image.shape[1:]), tuple(masks.shape[1:]), tuple(labels.shape[1:]), n)
        if self._key != key:
            torch.cuda.current_stream(self.device).synchronize()  # old buffers may still be in use
            self.copy_stream.synchronize()
            mk = lambda t: [torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device) for _ in range(2)]
            self._bufs = (mk(image), mk(masks), mk(labels))
            self._key = key
            self._n = 0
        return self._bufs

    def __call__(self, h_image, h_masks, h_labels, h_out=None, out_size=None, d_out=None):
        """Returns the pinned host label map [B,H,W] (uint8), or fills and returns the device tensor
        d_out instead when one is given (e.g. to all-gather the labels before the copy back)."""
        B = h_image.shape[0]
        H, W = _size_of(out_size) if out_size is not None else _size_of(h_image)
        if h_out is None and d_out is None:
            h_out = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
        nchunk = max(1, min(self.chunks, B))
        step = -(-B // nchunk)
        d_img, d_msk, d_lab = self._buffers(h_image, h_masks, h_labels, step)
        main = torch.cuda.current_stream(self.device)
        for lo in range(0, B, step):
            hi = min(lo + step, B)
            j, n = self._n & 1, hi - lo
            with torch.cuda.stream(self.copy_stream):
                if self._n >= 2:
                    self.copy_stream.wait_event(self._freed[j])  # set j was last used two chunks ago
                d_img[j][:n].copy_(h_image[lo:hi], non_blocking=True)
                d_msk[j][:n].copy_(h_masks[lo:hi], non_blocking=True)
                d_lab[j][:n].copy_(h_labels[lo:hi], non_blocking=True)
                self._ready[j].record(self.copy_stream)
            main.wait_event(self._ready[j])
            lab = refine_and_label(self.pamr, d_img[j][:n], d_msk[j][:n], d_lab[j][:n], (H, W))
            self._freed[j].record(main)
            (d_out if d_out is not None else h_out)[lo:hi].copy_(lab, non_blocking=True)
            self._n += 1
        return d_out if d_out is not None else h_out
