"""Golden vectors for SURVEY 8(f) row 4 -- the data set's denorm (datasets/pascal_voc.py:85-101, called at
train.py:120) followed by the image resize of run_pamr (models/SoftMaxAE.py:177) -- FROM THE REFERENCE ITSELF.
Build container only:   python oracle/gen_golden_denorm.py   -> tests/golden/denorm_*.npz"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("PAMR_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
from datasets.pascal_voc import PascalVOC  # noqa: E402  (reference, unmodified)

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    ds = PascalVOC.__new__(PascalVOC)  # denorm only reads the class constants MEAN / STD
    rng = np.random.RandomState(5)
    for (B, H, W, h, w) in [(2, 33, 41, 9, 11), (1, 24, 31, 24, 31), (2, 17, 13, 40, 33)]:
        image = rng.randn(B, 3, H, W).astype(np.float32)  # normalised network input
        with torch.no_grad():
            raw = ds.denorm(torch.from_numpy(image).clone())                                   # train.py:120
            small = F.interpolate(raw, (h, w), mode="bilinear", align_corners=True)            # SoftMaxAE.py:177
        name = "denorm_%dx%d_to_%dx%d" % (H, W, h, w)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), image=image, mean=np.float32(PascalVOC.MEAN),
                            std=np.float32(PascalVOC.STD), raw=raw.numpy(), out=small.numpy())
        print(name)


if __name__ == "__main__":
    main()
