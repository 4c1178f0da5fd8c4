"""PAMR nn.Module backed by libpamr_b200.so -- drop-in for the reference's models/mods/pamr.py.

Mirrors (reference paths relative to the reference repo root):
  PAMR(num_iter=1, dilations=[1]).forward(x, mask)      models/mods/pamr.py:114-143
  sub-modules aff_x / aff_m / aff_std with a `kernel`   models/mods/pamr.py:10-109
  buffer each, so state-dict keys, shapes and values are those of the reference and its
  snapshots load with strict=True (utils/checkpoints.py:99).

Differences, all deliberate:
  * CUDA fp32 only, sm_100 only, no CPU fallback: anything else raises RuntimeError;
  * forward-only: the result never carries autograd history (every reference call site
    detaches its inputs, models/SoftMaxAE.py:251,258);
  * the per-call device-synchronising tamper assert (pamr.py:42-43) is replaced by a host-side
    check of the constant buffers at construction and after load_state_dict.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib


def _shift_kernel(centre, neighbour, with_centre_tap):
    """[8 or 9, 1, 3, 3] fixed kernel: one 3x3 position per output channel in row-major order
    (the centre position is skipped unless with_centre_tap), value `neighbour` there and
    `centre` at the middle."""
    taps = [(r, c) for r in range(3) for c in range(3) if with_centre_tap or (r, c) != (1, 1)]
    k = torch.zeros(len(taps), 1, 3, 3)
    for i, (r, c) in enumerate(taps):
        k[i, 0, 1, 1] = centre
        k[i, 0, r, c] = neighbour
    return k


class LocalAffinity(nn.Module):
    """Holder of the fixed difference kernel (reference pamr.py:10-55): centre +1, neighbour -1.
    The unfolding itself happens inside the CUDA kernels; this module only owns the buffer."""

    def __init__(self, dilations=[1]):
        super().__init__()
        self.dilations = list(dilations)
        weight = self._init_aff()
        self.register_buffer("kernel", weight)
        self.weight_check = weight.clone()

    def _init_aff(self):
        return _shift_kernel(1.0, -1.0, False)

    def check_kernel(self):
        """Host-side replacement of `assert torch.all(weight_check.eq(kernel))` (pamr.py:42-43)."""
        assert torch.equal(self.kernel.detach().cpu().float(), self.weight_check), \
            "%s.kernel was altered" % type(self).__name__

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.check_kernel()

    def forward(self, x):
        raise RuntimeError("the neighbour unfolding is fused into libpamr_b200's kernels; call PAMR.forward")


class LocalAffinityCopy(LocalAffinity):
    """pamr.py:57-75: pure shift (+1 at the neighbour)."""

    def _init_aff(self):
        return _shift_kernel(0.0, 1.0, False)


class LocalStDev(LocalAffinity):
    """pamr.py:77-103: 9 copy taps including the centre."""

    def _init_aff(self):
        return _shift_kernel(0.0, 1.0, True)

    def forward(self, x):
        """x [B,K,H,W] -> unbiased std over the 9*nd samples, [B,K,1,H,W] like the reference (pamr.py:98-103)."""
        return local_std(x, self.dilations).unsqueeze(2)


class LocalAffinityAbs(LocalAffinity):
    """pamr.py:105-109: |centre - neighbour|."""


def _dil_array(dilations):
    d = [int(v) for v in dilations]
    return (ctypes.c_int * len(d))(*d), len(d)


def _check_cuda_f32(name, t, ndim=4):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: libpamr_b200 runs on CUDA (sm_100) only and has no CPU fallback"
                           % (name, t.device))
    if t.dtype != torch.float32:
        raise RuntimeError("%s has dtype %s: libpamr_b200 computes in float32 only" % (name, t.dtype))
    if t.dim() != ndim:
        raise RuntimeError("%s must be %d-dimensional, got shape %s" % (name, ndim, tuple(t.shape)))
    return t.detach().contiguous()


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dev(t):
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def resize_bilinear(x, size):
    """F.interpolate(x, size, mode='bilinear', align_corners=True) on the GPU (pamr.py:125)."""
    x = _check_cuda_f32("x", x)
    B, Ch, h, w = x.shape
    H, W = int(size[0]), int(size[1])
    out = torch.empty((B, Ch, H, W), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().pamr_resize_bilinear_f32(x.data_ptr(), out.data_ptr(), B * Ch, h, w, H, W, _dev(x),
                                                   _stream(x.device)))
    return out


def local_affinity(x, dilations):
    """pamr.py:132-136: image [B,K,H,W] -> softmax affinity [B,8*nd,H,W]."""
    x = _check_cuda_f32("x", x)
    B, K, H, W = x.shape
    d, nd = _dil_array(dilations)
    aff = torch.empty((B, 8 * nd, H, W), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().pamr_affinity_f32(x.data_ptr(), aff.data_ptr(), B, K, H, W, d, nd, _dev(x), _stream(x.device)))
    return aff


def local_std(x, dilations):
    """LocalStDev (pamr.py:77-103): image [B,K,H,W] -> unbiased std over the 9*nd samples, [B,K,H,W]."""
    x = _check_cuda_f32("x", x)
    B, K, H, W = x.shape
    d, nd = _dil_array(dilations)
    sd = torch.empty_like(x)
    _lib.check(_lib.lib().pamr_local_std_f32(x.data_ptr(), sd.data_ptr(), B, K, H, W, d, nd, _dev(x), _stream(x.device)))
    return sd


def propagate(aff, mask, dilations, num_iter, return_class_max=False):
    """pamr.py:138-140: num_iter affinity-weighted neighbour averages of mask [B,C,H,W]."""
    aff = _check_cuda_f32("aff", aff)
    mask = _check_cuda_f32("mask", mask)
    B, C, H, W = mask.shape
    d, nd = _dil_array(dilations)
    if tuple(aff.shape) != (B, 8 * nd, H, W):
        raise RuntimeError("aff has shape %s, expected %s" % (tuple(aff.shape), (B, 8 * nd, H, W)))
    out = torch.empty_like(mask)
    L = _lib.lib()
    nbytes = L.pamr_propagate_scratch_bytes(B, C, H, W, d, nd, int(num_iter))
    tmp = torch.empty((nbytes,), dtype=torch.uint8, device=mask.device) if nbytes else None
    cmax = torch.empty((B, C), dtype=torch.int32, device=mask.device) if return_class_max else None
    _lib.check(L.pamr_propagate_f32(
        aff.data_ptr(), mask.data_ptr(), out.data_ptr(), tmp.data_ptr() if tmp is not None else None, nbytes, B, C, H, W,
        d, nd, int(num_iter), cmax.data_ptr() if cmax is not None else None, _dev(mask), _stream(mask.device)))
    return (out, cmax) if return_class_max else out


class PAMR(nn.Module):
    """Pixel-adaptive mask refinement; same constructor, attributes and state dict as the reference."""

    def __init__(self, num_iter=1, dilations=[1]):
        super().__init__()
        self.num_iter = num_iter
        self.aff_x = LocalAffinityAbs(dilations)
        self.aff_m = LocalAffinityCopy(dilations)
        self.aff_std = LocalStDev(dilations)

    @property
    def dilations(self):
        return self.aff_x.dilations

    def forward(self, x, mask, return_class_max=False):
        """x: image [B,K,H,W]; mask: [B,C,h,w] (resized to [H,W] first, pamr.py:125) -> [B,C,H,W].

        With return_class_max=True also returns the per-(b,c) maximum of the result in the
        library's ordered-uint encoding (int32 tensor [B,C]) for stage.pseudo_labels."""
        x = _check_cuda_f32("x", x)
        mask = _check_cuda_f32("mask", mask)
        if mask.device != x.device:
            raise RuntimeError("x is on %s but mask is on %s" % (x.device, mask.device))
        B, K, H, W = x.shape
        Bm, C, h, w = mask.shape
        if Bm != B:
            raise RuntimeError("batch size mismatch: x has %d, mask has %d" % (B, Bm))
        d, nd = _dil_array(self.dilations)
        L = _lib.lib()
        iters = int(self.num_iter)
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
        nbytes = L.pamr_forward_workspace_bytes(B, K, C, H, W, h, w, d, nd, iters)
        # torch's caching allocator returns >= 512-byte aligned blocks
        ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=x.device)
        cmax = torch.empty((B, C), dtype=torch.int32, device=x.device) if return_class_max else None
        _lib.check(L.pamr_forward_f32(x.data_ptr(), mask.data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes, B, K, C, H,
                                      W, h, w, d, nd, iters, cmax.data_ptr() if cmax is not None else None, _dev(x),
                                      _stream(x.device)))
        return (out, cmax) if return_class_max else out
