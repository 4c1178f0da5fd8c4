"""Generate golden vectors for the PAMR hot path FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):   python oracle/gen_golden.py
It imports the reference's unmodified modules
  models/mods/pamr.py        (PAMR, LocalStDev, LocalAffinityAbs)
  models/SoftMaxAE.py        (run_pamr, _rescale_and_clean, pseudo_gtmask, argmax lines 62-67)
on CPU fp32 and writes inputs + outputs of small cases to tests/golden/*.npz.
The oracle (oracle/pamr_oracle.c) and the CUDA path are both tested against these files.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("PAMR_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402
from models.mods.pamr import PAMR  # noqa: E402  (reference)
from models.SoftMaxAE import network_SoftMaxAE, pseudo_gtmask  # noqa: E402  (reference)

torch.set_grad_enabled(False)
OUT = os.path.join(ROOT, "tests", "golden")
D6 = [1, 2, 4, 8, 12, 24]


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def ref_intermediates(pamr, x):
    """std and softmax affinity exactly as PAMR.forward computes them (pamr.py:132-136)."""
    x_std = pamr.aff_std(x)
    a = -pamr.aff_x(x) / (1e-8 + 0.1 * x_std)
    a = a.mean(1, keepdim=True)
    a = F.softmax(a, 2)
    return x_std[:, :, 0].numpy(), a[:, 0].numpy()


def ref_labels(pseudo_gt):
    """SoftMaxAE.py:62-67 verbatim semantics."""
    mask_gt = torch.argmax(pseudo_gt, 1)
    ignore_mask = pseudo_gt.sum(1) < 1.
    mask_gt[ignore_mask] = 255
    return mask_gt.numpy().astype(np.uint8)


def save(name, **kw):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print("%-28s %7.1f KB" % (name, os.path.getsize(path) / 1024))


def case_pamr(name, image, mask, num_iter=10, dilations=D6, every_iter=False):
    pamr = PAMR(num_iter, dilations)
    out = pamr(T(image), T(mask)).numpy()
    kw = dict(image=image, mask=mask, num_iter=num_iter, dilations=np.array(dilations), out=out)
    if image.shape[-2:] == mask.shape[-2:] or True:
        sd, aff = ref_intermediates(pamr, T(image))
        kw.update(std=sd, aff=aff)
    if every_iter:
        kw["out_iter1"] = PAMR(1, dilations)(T(image), T(mask)).numpy()
    save(name, **kw)


def main():
    os.makedirs(OUT, exist_ok=True)
    Net = network_SoftMaxAE(types.SimpleNamespace(BACKBONE="resnet38"))

    # 1. uniform image, all dilations
    case_pamr("pamr_uniform_37x29", synth.image_uniform(2, 3, 37, 29, 0), synth.mask_softmax(2, 5, 37, 29, 10),
              every_iter=True)
    # 2. structured / quantised image, 21 classes, mask at lower resolution (upsampled at pamr.py:125)
    case_pamr("pamr_quant_48x64_lowres", synth.image_structured(1, 3, 48, 64, 2, quantise=True),
              synth.mask_blobs(1, 21, 12, 16, 11))
    case_pamr("pamr_struct_40x56", synth.image_structured(1, 3, 40, 56, 1), synth.mask_softmax(1, 21, 40, 56, 12))
    # 3. constant image: sigma == 0
    case_pamr("pamr_const_30x30", synth.image_constant(1, 3, 30, 30), synth.mask_softmax(1, 4, 30, 30, 13))
    # 4. edge shapes (H, W below / at / just above the largest dilation)
    for (h, w) in [(1, 1), (1, 7), (8, 8), (24, 24), (25, 25), (5, 50)]:
        case_pamr("pamr_edge_%dx%d" % (h, w), synth.image_uniform(1, 3, h, w, 3), synth.mask_softmax(1, 3, h, w, 14))
    # 5. non-default constructor arguments: K=1 channel, dilations [1,3], 3 iterations; K=4, single dilation
    case_pamr("pamr_k1_d13_it3", synth.image_uniform(2, 1, 19, 23, 4), synth.mask_softmax(2, 6, 19, 23, 15),
              num_iter=3, dilations=[1, 3])
    case_pamr("pamr_k4_d5_it1", synth.image_structured(1, 4, 16, 33, 5), synth.mask_softmax(1, 2, 16, 33, 16),
              num_iter=1, dilations=[5])

    # 6. bilinear align_corners=True resizes (SoftMaxAE.py:177, :266): integer and non-integer ratios
    for (h, w, H, W) in [(21, 21, 81, 81), (13, 17, 40, 33), (40, 33, 13, 17), (1, 5, 4, 9), (7, 7, 7, 7), (6, 9, 1, 1)]:
        x = synth.image_uniform(2, 3, h, w, 6)
        y = F.interpolate(T(x), size=(H, W), mode="bilinear", align_corners=True).numpy()
        save("resize_%dx%d_to_%dx%d" % (h, w, H, W), x=x, y=y)

    # 7. stage_net sequence A (SoftMaxAE.py:250-259): PAMR(raw softmax) -> clean -> pseudo_gt -> labels
    B, C, h, w, H, W = 2, 21, 21, 25, 81, 97
    image = synth.image_structured(B, 3, H, W, 7)
    masks = synth.mask_blobs(B, C, h, w, 17)
    labels = synth.labels_bernoulli(B, C, 20, p=0.3)
    stub = types.SimpleNamespace(_aff=PAMR(10, D6))
    masks_dec = Net.run_pamr(stub, T(image), T(masks))
    cleaned = Net._rescale_and_clean(stub, masks_dec, T(image), T(labels))
    pg = pseudo_gtmask(cleaned)
    save("stage_seqA_21x25_to_81x97", image=image, masks=masks, labels=labels, masks_dec=masks_dec.numpy(),
         cleaned=cleaned.numpy(), pseudo_gt=pg.numpy().astype(np.uint8), label=ref_labels(pg))

    # 8. stage_net sequence B (CAM_CASA_WGAP_tf.py:335-345): clean@mask-res -> PAMR -> clean@image-res
    pre = Net._rescale_and_clean(stub, T(masks), T(masks), T(labels))
    masks_dec_b = Net.run_pamr(stub, T(image), pre)
    cleaned_b = Net._rescale_and_clean(stub, masks_dec_b, T(image), T(labels))
    pg_b = pseudo_gtmask(cleaned_b)
    save("stage_seqB_21x25_to_81x97", image=image, masks=masks, labels=labels, pre=pre.numpy(),
         masks_dec=masks_dec_b.numpy(), cleaned=cleaned_b.numpy(), pseudo_gt=pg_b.numpy().astype(np.uint8),
         label=ref_labels(pg_b))

    # 9. benchmark-style call: mask already at image resolution, 21 classes (config-2 shape family, small)
    B, C, H, W = 2, 21, 65, 77
    image = synth.image_structured(B, 3, H, W, 8, quantise=True)
    masks = synth.mask_blobs(B, C, H, W, 18)
    labels = synth.labels_bernoulli(B, C, 21, p=0.3)
    masks_dec = Net.run_pamr(stub, T(image), T(masks))
    cleaned = Net._rescale_and_clean(stub, masks_dec, T(image), T(labels))
    pg = pseudo_gtmask(cleaned)
    save("stage_fullres_65x77", image=image, masks=masks, labels=labels, masks_dec=masks_dec.numpy(),
         cleaned=cleaned.numpy(), pseudo_gt=pg.numpy().astype(np.uint8), label=ref_labels(pg))

    # 10. state-dict keys / shapes / values of the reference module (drop-in contract)
    sd = PAMR(10, D6).state_dict()
    save("state_dict", **{k: v.numpy() for k, v in sd.items()})


if __name__ == "__main__":
    main()
