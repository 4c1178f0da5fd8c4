"""numpy/ctypes front-end of the CPU oracle (oracle/pamr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Parity pinning: checked against golden vectors generated from the reference's own
Python modules (oracle/gen_golden.py -> tests/golden/*.npz) in tests/test_oracle_golden.py.

Function names follow the reference: PAMR.forward (models/mods/pamr.py:124-143),
run_pamr / _rescale_and_clean / pseudo_gtmask (models/SoftMaxAE.py:176-179, 263-268, 29-50).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpamr_oracle.so")
_lib = None

DEFAULT_DILATIONS = (1, 2, 4, 8, 12, 24)  # core/config.py:92
DEFAULT_ITERS = 10  # core/config.py:93

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)
_u8 = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """Compile oracle/pamr_oracle.c with gcc (building the checker is not using it)."""
    src = os.path.join(_HERE, "pamr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libpamr_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.pamr_oracle_num_threads.restype = ctypes.c_int
        L.pamr_oracle_set_num_threads.argtypes = [ctypes.c_int]
        L.pamr_oracle_local_std.argtypes = [_f, _f] + [ctypes.c_int] * 4 + [_i, ctypes.c_int]
        L.pamr_oracle_affinity.argtypes = [_f, _f] + [ctypes.c_int] * 4 + [_i, ctypes.c_int]
        L.pamr_oracle_propagate.argtypes = [_f, _f, _f] + [ctypes.c_int] * 4 + [_i, ctypes.c_int, ctypes.c_int]
        L.pamr_oracle_resize_bilinear.argtypes = [_f, _f] + [ctypes.c_int] * 5
        L.pamr_oracle_gate.argtypes = [_f, _f] + [ctypes.c_int] * 4
        L.pamr_oracle_pseudo_gt.argtypes = [_f, _f, _u8] + [ctypes.c_int] * 4 + [ctypes.c_float] * 3
        L.pamr_oracle_forward.argtypes = [_f, _f, _f] + [ctypes.c_int] * 7 + [_i, ctypes.c_int, ctypes.c_int]
        L.pamr_oracle_merge_multiscale.argtypes = [_f, _i, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p] + \
            [ctypes.c_int] * 7 + [ctypes.c_float] * 2
        L.pamr_oracle_mask_ce.argtypes = [_f, _f, _f, _f, ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 6
        _lib = L
    return _lib


def num_threads():
    return lib().pamr_oracle_num_threads()


def set_num_threads(n):
    lib().pamr_oracle_set_num_threads(int(n))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(_f)


def _dil(dilations):
    d = np.ascontiguousarray(list(dilations), dtype=np.int32)
    return d, d.ctypes.data_as(_i), len(d)


def local_std(image, dilations=DEFAULT_DILATIONS):
    """LocalStDev.forward (pamr.py:98-103): [B,K,H,W] -> [B,K,H,W] unbiased std over 9*nd samples."""
    image = _c(image)
    B, K, H, W = image.shape
    out = np.empty_like(image)
    d, dp, nd = _dil(dilations)
    lib().pamr_oracle_local_std(_p(image), _p(out), B, K, H, W, dp, nd)
    return out


def affinity(image, dilations=DEFAULT_DILATIONS):
    """pamr.py:132-136: [B,K,H,W] -> softmax affinity [B,8*nd,H,W]."""
    image = _c(image)
    B, K, H, W = image.shape
    d, dp, nd = _dil(dilations)
    out = np.empty((B, 8 * nd, H, W), dtype=np.float32)
    lib().pamr_oracle_affinity(_p(image), _p(out), B, K, H, W, dp, nd)
    return out


def propagate(aff, mask, dilations=DEFAULT_DILATIONS, num_iter=DEFAULT_ITERS):
    """pamr.py:138-140: num_iter affinity-weighted neighbour averages."""
    aff, mask = _c(aff), _c(mask)
    B, C, H, W = mask.shape
    d, dp, nd = _dil(dilations)
    assert aff.shape == (B, 8 * nd, H, W)
    out = np.empty_like(mask)
    lib().pamr_oracle_propagate(_p(aff), _p(mask), _p(out), B, C, H, W, dp, nd, int(num_iter))
    return out


def resize_bilinear(x, size):
    """F.interpolate(x, size, mode='bilinear', align_corners=True) on [B,Ch,h,w]."""
    x = _c(x)
    B, Ch, h, w = x.shape
    H, W = int(size[0]), int(size[1])
    out = np.empty((B, Ch, H, W), dtype=np.float32)
    lib().pamr_oracle_resize_bilinear(_p(x), _p(out), B * Ch, h, w, H, W)
    return out


def pamr_forward(image, mask, num_iter=DEFAULT_ITERS, dilations=DEFAULT_DILATIONS):
    """PAMR(num_iter, dilations).forward(image, mask) (pamr.py:124-143)."""
    image, mask = _c(image), _c(mask)
    B, K, H, W = image.shape
    _, C, h, w = mask.shape
    d, dp, nd = _dil(dilations)
    out = np.empty((B, C, H, W), dtype=np.float32)
    lib().pamr_oracle_forward(_p(image), _p(mask), _p(out), B, K, C, H, W, h, w, dp, nd, int(num_iter))
    return out


def run_pamr(image, mask, num_iter=DEFAULT_ITERS, dilations=DEFAULT_DILATIONS):
    """run_pamr (SoftMaxAE.py:176-179): image is resized to the mask size, then PAMR."""
    im = resize_bilinear(image, mask.shape[-2:])
    return pamr_forward(im, mask, num_iter, dilations)


def rescale_and_clean(masks, size, labels):
    """_rescale_and_clean (SoftMaxAE.py:263-268): bilinear to `size`, then masks[:,1:] *= labels."""
    out = resize_bilinear(masks, size)
    B, C, H, W = out.shape
    labels = _c(labels)
    assert labels.shape == (B, C - 1)
    lib().pamr_oracle_gate(_p(out), _p(labels), B, C, H, W)
    return out


def pseudo_gtmask(mask, cutoff_top=0.6, cutoff_low=0.2, cutoff_bg=0.7):
    """pseudo_gtmask (SoftMaxAE.py:29-50) -> one-hot-or-empty float [B,C,H,W]."""
    mask = _c(mask)
    B, C, H, W = mask.shape
    pg = np.empty_like(mask)
    lib().pamr_oracle_pseudo_gt(_p(mask), _p(pg), None, B, C, H, W, cutoff_bg, cutoff_top, cutoff_low)
    return pg


def pseudo_labels(mask, cutoff_top=0.6, cutoff_low=0.2, cutoff_bg=0.7):
    """pseudo_gtmask followed by argmax / ignore-255 (SoftMaxAE.py:61-67) -> uint8 [B,H,W]."""
    mask = _c(mask)
    B, C, H, W = mask.shape
    lab = np.empty((B, H, W), dtype=np.uint8)
    lib().pamr_oracle_pseudo_gt(_p(mask), None, lab.ctypes.data_as(_u8), B, C, H, W, cutoff_bg, cutoff_top, cutoff_low)
    return lab


def thresholds(mask, cutoff_top=0.6, cutoff_low=0.2, cutoff_bg=0.7):
    """Per-(b,c) thresholds of pseudo_gtmask (SoftMaxAE.py:35-42), float32 arithmetic."""
    mask = _c(mask)
    B, C = mask.shape[:2]
    mx = mask.reshape(B, C, -1).max(-1)
    cut = np.full((1, C), np.float32(cutoff_top), dtype=np.float32)
    cut[0, 0] = np.float32(cutoff_bg)
    return np.maximum((mx * cut).astype(np.float32), np.float32(cutoff_low))


def near_threshold_set(mask, tol=2e-5, **kw):
    """SURVEY 8(a) label-parity rule: pixels where some class is within tol of its threshold."""
    thr = thresholds(mask, **kw)
    return (np.abs(_c(mask) - thr[:, :, None, None]).min(1) <= tol)


def balanced_mask_loss_ce(logits, pseudo_gt, gt_labels, gout=None):
    """balanced_mask_loss_ce (SoftMaxAE.py:52-88): loss [B]; with gout [B] also the gradient of
    sum_b gout[b]*loss[b] w.r.t. logits.  pseudo_gt is the float one-hot-or-empty tensor [B,C,H,W]."""
    logits, pg, gl = _c(logits), _c(np.asarray(pseudo_gt, dtype=np.float32)), _c(gt_labels)
    B, C, h, w = logits.shape
    H, W = pg.shape[-2:]
    loss = np.empty((B,), dtype=np.float32)
    if gout is None:
        lib().pamr_oracle_mask_ce(_p(logits), _p(pg), _p(gl), _p(loss), None, None, B, C, h, w, H, W)
        return loss
    gout = _c(gout)
    grad = np.empty_like(logits)
    lib().pamr_oracle_mask_ce(_p(logits), _p(pg), _p(gl), _p(loss), gout.ctypes.data_as(ctypes.c_void_p),
                              grad.ctypes.data_as(ctypes.c_void_p), B, C, h, w, H, W)
    return loss, grad


def merge_multiscale(masks, pads, labels, imsize_hw, flip, bg_pow, prospect_thresh):
    """MergeMultiScale._merge_masks (utils/inference_tools.py:134-161) + the no-CRF prediction of
    ResultWriter.save (:85-88): returns (merged [C,H,W] float32, pred [H,W] uint8)."""
    masks = _c(masks)
    S, C, Hp, Wp = masks.shape
    H, W = int(imsize_hw[0]), int(imsize_hw[1])
    pads = np.ascontiguousarray(pads, dtype=np.int32)
    lab = None if labels is None else _c(labels)
    merged = np.empty((C, H, W), dtype=np.float32)
    pred = np.empty((H, W), dtype=np.uint8)
    lib().pamr_oracle_merge_multiscale(_p(masks), pads.ctypes.data_as(_i), None if lab is None else lab.ctypes.data_as(ctypes.c_void_p),
                                       merged.ctypes.data_as(ctypes.c_void_p), pred.ctypes.data_as(ctypes.c_void_p), S, C, Hp, Wp,
                                       H, W, int(bool(flip)), float(bg_pow), float(prospect_thresh))
    return merged, pred


def denorm_resize(image_norm, mean, std, size=None):
    """denorm (datasets/pascal_voc.py:85-101: t.mul_(s).add_(m) per channel, two float roundings) followed by
    the align_corners=True bilinear resize of run_pamr (SoftMaxAE.py:177)."""
    x = _c(image_norm).copy()
    for k, (m, s_) in enumerate(zip(mean, std)):
        x[:, k] = x[:, k] * np.float32(s_) + np.float32(m)
    return x if size is None or tuple(size) == x.shape[-2:] else resize_bilinear(x, size)
