"""Probe: does the affinity tile kernel (1-D TMA rows at 4-byte granular starts) work for W not a multiple of 4?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synth, wseg_b200
from oracle import oracle
D6 = [1, 2, 4, 8, 12, 24]
W = int(sys.argv[1]); H = int(sys.argv[2]) if len(sys.argv) > 2 else 130; B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
image, mask = synth.image_structured(B, 3, H, W, 5), synth.mask_softmax(B, 21, H, W, 6)
G = lambda a: torch.from_numpy(a).to("cuda:0")
out = wseg_b200.PAMR(1, D6).to("cuda:0")(G(image), G(mask))
torch.cuda.synchronize()
ref = oracle.pamr_forward(image, mask, 1, D6)
print("B=%d " % B + "W=%d H=%d: max-abs %.3g" % (W, H, np.abs(out.cpu().numpy() - ref).max()), flush=True)
