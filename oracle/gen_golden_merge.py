"""Golden vectors for SURVEY 8(f) row 3 -- inference post-processing
  MergeMultiScale._merge_masks   (utils/inference_tools.py:134-161)   un-pad, resize, un-flip, gate, mean, BG_POW
  ResultWriter.save, no-CRF path (utils/inference_tools.py:85-88)     fg < prospect_thresh -> 0, argmax
FROM THE REFERENCE ITSELF (CPU fp32).  Build container only:   python oracle/gen_golden_merge.py

utils/inference_tools.py imports matplotlib, pydensecrf (utils/dcrf) and the data-set loaders at module level;
none of them is used by the two code paths above, so they are stubbed in sys.modules and the reference's class is
imported UNMODIFIED.  `save` itself writes PNG files through scipy.misc.imsave (gone from current scipy), so its
four post-processing lines (:85-88) are restated here verbatim (ref_predict)."""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("PAMR_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "tests"))
for name in ["matplotlib", "matplotlib.pyplot", "utils.dcrf", "datasets.pascal_voc_ms", "scipy.misc"]:
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["utils.dcrf"].crf_inference = None
sys.modules["datasets.pascal_voc_ms"].MultiscaleLoader = None
sys.modules["datasets.pascal_voc_ms"].CropLoader = None
import scipy  # noqa: E402
scipy.misc = sys.modules["scipy.misc"]

import synth  # noqa: E402
from utils.inference_tools import MergeMultiScale  # noqa: E402  (reference, unmodified)

OUT = os.path.join(ROOT, "tests", "golden")


def ref_predict(merged_mask, prospect_thresh):
    """utils/inference_tools.py:85-88 verbatim semantics."""
    merged_mask = merged_mask.copy()
    index = list(np.where(merged_mask[1:, :, :] < prospect_thresh))
    index[0] += 1
    merged_mask[tuple(index)] = 0
    return np.argmax(merged_mask, 0).astype(np.uint8)


def case(name, S, C, pad_hw, im_hw, scales, flip, bg_pow, thresh, seed):
    rng = np.random.RandomState(seed)
    Hp, Wp = pad_hw
    H, W = im_hw
    masks = np.zeros((S, C, Hp, Wp), dtype=np.float32)
    pads = np.zeros((S, 4), dtype=np.float32)
    for s in range(S):
        sc = scales[s // (2 if flip else 1)]
        h, w = int(round(H * sc)), int(round(W * sc))
        pt, pl = (Hp - h) // 2, (Wp - w) // 2
        m = synth.mask_softmax(1, C, h, w, seed * 10 + s)[0]
        masks[s, :, pt:pt + h, pl:pl + w] = m
        pads[s] = (pt, pl, h, w)
    labels = (rng.rand(C - 1) < 0.4).astype(np.float32)
    labels[0] = 1.0
    cfg = types.SimpleNamespace(FLIP=flip, BG_POW=bg_pow)
    writer = MergeMultiScale(cfg, None, "/tmp", prospect_thresh=thresh, verbose=False, heatmap=False, scoremap=False, CRF=False)
    with torch.no_grad():
        merged = writer._merge_masks(torch.from_numpy(masks.copy()), torch.from_numpy(labels), torch.from_numpy(pads), (H, W))
    pred = ref_predict(merged, thresh)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, masks=masks, pads=pads.astype(np.int32), labels=labels, flip=np.int32(flip), bg_pow=np.float32(bg_pow),
                        thresh=np.float32(thresh), merged=merged.astype(np.float32), pred=pred)
    print("%-32s %7.1f KB  merged dtype %s  classes used %s" % (name, os.path.getsize(path) / 1024, merged.dtype, np.unique(pred)))


def main():
    # the shipped configuration family: 4 scales x flip, BG_POW = 3 (core/config.py:44-50), small sizes
    case("merge_8x5_flip_pow3", 8, 5, (64, 88), (30, 41), [1, 0.5, 1.5, 2.0], True, 3, 0.3, 1)
    case("merge_4x21_noflip_pow3", 4, 21, (48, 48), (23, 19), [1, 0.75, 1.25, 1.5], False, 3, 0.05, 2)
    case("merge_2x4_flip_pow1_thr0", 2, 4, (20, 20), (17, 13), [1], True, 1, 0.0, 3)
    case("merge_1x3_single", 1, 3, (9, 12), (9, 12), [1], False, 3, 0.7, 4)


if __name__ == "__main__":
    main()
