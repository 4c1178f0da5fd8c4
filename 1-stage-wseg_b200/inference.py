"""Inference post-processing, backed by libpamr_b200.so.

Mirrors, for one image,
  MergeMultiScale._merge_masks(masks, labels, pads, imsize_hw)   utils/inference_tools.py:134-161
  the no-CRF prediction of ResultWriter.save                     utils/inference_tools.py:85-88
on the GPU: the reference moves the [S,C,Hp,Wp] network output to the host (infer_val.py:124) and merges it
with numpy; here the scores stay on the device and only the uint8 prediction needs to travel.
"""
import ctypes

import torch

from . import _lib
from .pamr import _check_cuda_f32, _dev, _stream


def _call(masks, labels, pads, imsize_hw, flip, bg_pow, prospect_thresh, want_merged, want_pred):
    masks = _check_cuda_f32("masks", masks)
    S, C, Hp, Wp = masks.shape
    H, W = int(imsize_hw[0]), int(imsize_hw[1])
    pads = [int(v) for row in (pads.tolist() if isinstance(pads, torch.Tensor) else pads) for v in row]
    if len(pads) != 4 * S:
        raise RuntimeError("pads must hold (pad_t, pad_l, h, w) for each of the %d scales" % S)
    c_pads = (ctypes.c_int * (4 * S))(*pads)
    lab = None
    if labels is not None:
        lab = labels.detach().to(device=masks.device, dtype=torch.float32).contiguous()
        if tuple(lab.shape) != (C - 1,):
            raise RuntimeError("labels must have shape [C-1] = (%d,), got %s" % (C - 1, tuple(lab.shape)))
    merged = torch.empty((C, H, W), dtype=torch.float32, device=masks.device) if want_merged else None
    pred = torch.empty((H, W), dtype=torch.uint8, device=masks.device) if want_pred else None
    _lib.check(_lib.lib().pamr_merge_multiscale_f32(
        masks.data_ptr(), ctypes.cast(c_pads, ctypes.c_void_p), lab.data_ptr() if lab is not None else None,
        merged.data_ptr() if merged is not None else None, pred.data_ptr() if pred is not None else None, S, C, Hp, Wp, H, W,
        int(bool(flip)), float(bg_pow), float(prospect_thresh), _dev(masks), _stream(masks.device)))
    return merged, pred


def merge_masks(masks, labels, pads, imsize_hw, flip=False, bg_pow=3):
    """MergeMultiScale._merge_masks (utils/inference_tools.py:134-161): masks [S,C,Hp,Wp], labels [C-1],
    pads [S,4] = (pad_t, pad_l, h, w); cfg.FLIP / cfg.BG_POW become arguments.  Returns [C,H,W]."""
    return _call(masks, labels, pads, imsize_hw, flip, bg_pow, 0.0, True, False)[0]


def merge_and_predict(masks, labels, pads, imsize_hw, prospect_thresh, flip=False, bg_pow=3, return_merged=False):
    """_merge_masks followed by the thresholded argmax of ResultWriter.save (:85-88), in one kernel.
    Returns the uint8 prediction [H,W] (and the merged scores [C,H,W] on request)."""
    merged, pred = _call(masks, labels, pads, imsize_hw, flip, bg_pow, prospect_thresh, return_merged, True)
    return (pred, merged) if return_merged else pred
