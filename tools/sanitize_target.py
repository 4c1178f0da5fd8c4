"""Smallest end-to-end invocation (tuned kernel + strips + epilogue) for compute-sanitizer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, synth, wseg_b200
dev = "cuda:0"
B, C, H, W = 1, 4, 41, 65
image, mask = synth.image_structured(B, 3, H, W, 1), synth.mask_blobs(B, C, H, W, 2)
labels = synth.labels_bernoulli(B, C, 3, p=0.5)
pamr = wseg_b200.PAMR(2, [1, 2, 4, 8, 12, 24]).to(dev)
G = lambda a: torch.from_numpy(a).to(dev)
lab = wseg_b200.refine_and_label(pamr, G(image), G(mask), G(labels))
torch.cuda.synchronize()
print("ok", lab.shape, int(lab.sum()))
