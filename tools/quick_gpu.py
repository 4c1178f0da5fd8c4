"""Quick GPU check of the tuned propagation kernel against the generic one and the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synth, wseg_b200
from oracle import oracle
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
G = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for (B, C, H, W) in [(1, 3, 40, 64), (2, 21, 64, 80), (1, 21, 97, 129), (1, 21, 321, 321), (2, 5, 100, 40)]:
    image, mask = synth.image_structured(B, 3, H, W, 5), synth.mask_softmax(B, C, H, W, 6)
    aff = oracle.affinity(image, D6)
    ref1 = oracle.propagate(aff, mask, D6, 1)
    out1, cm = wseg_b200.propagate(G(aff), G(mask), D6, 1, return_class_max=True)
    torch.cuda.synchronize()
    e1 = np.abs(out1.cpu().numpy() - ref1).max()
    ref = oracle.pamr_forward(image, mask, 10, D6)
    out = wseg_b200.PAMR(10, D6).to(dev)(G(image), G(mask))
    torch.cuda.synchronize()
    print((B, C, H, W), "1 iter err %.3g" % e1, "10 iter err %.3g" % np.abs(out.cpu().numpy() - ref).max(), flush=True)
B, C, H, W = 16, 21, 321, 321
image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, H, W), device=dev), 1)
aff = wseg_b200.local_affinity(image, D6)
for _ in range(3): wseg_b200.propagate(aff, mask, D6, 10)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): wseg_b200.propagate(aff, mask, D6, 10)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("propagate x10: %.3f ms  -> %.3f ms/iter, %.1f GB/s algorithmic (%.1f%% of 6543)" % (ms, ms / 10, 360 * B * H * W / (ms / 10 * 1e-3) / 1e9, 360 * B * H * W / (ms / 10 * 1e-3) / 1e9 / 65.431))
