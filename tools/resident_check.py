"""Small-map resident kernel: parity against the oracle on a few shapes + timing of PAMR.forward at the
stage_net call shapes (B=16, 41x41 / 81x81).  PAMR_B200_NO_RESIDENT=1 (experiment builds) times the tiled path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synth, wseg_b200
from oracle import oracle
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
G = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
if os.environ.get("RC_SKIP_PARITY") != "1":
    for (B, K, C, H, W, it) in [(2, 3, 21, 41, 41, 10), (1, 3, 21, 81, 81, 10), (3, 3, 5, 33, 57, 3), (2, 1, 4, 20, 100, 1),
                                (16, 3, 21, 41, 41, 10), (2, 5, 7, 64, 64, 2), (1, 3, 21, 7, 9, 10), (150, 3, 2, 12, 12, 2),
                                (40, 3, 21, 41, 41, 10)]:
        image, mask = synth.image_structured(B, K, H, W, 5), synth.mask_softmax(B, C, H, W, 6)
        ref = oracle.pamr_forward(image, mask, it, D6)
        n0 = wseg_b200._lib.launch_count()
        out, cm = wseg_b200.PAMR(it, D6).to(dev)(G(image), G(mask), return_class_max=True)
        torch.cuda.synchronize()
        nl = wseg_b200._lib.launch_count() - n0
        o = out.cpu().numpy()
        mx = np.array([[wseg_b200._lib.lib().pamr_float_from_ordered(int(v) & 0xffffffff) for v in row] for row in cm.cpu().numpy()], np.float32)
        print((B, K, C, H, W, it), "err %.3g" % np.abs(o - ref).max(), "class max exact:", bool((mx == o.reshape(B, C, -1).max(-1)).all()),
              "launches", nl, flush=True)
def timed(pamr, image, mask, n=20):
    for _ in range(3): pamr(image, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): pamr(image, mask)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (B, H, W) in [(16, 41, 41), (16, 81, 81), (1, 81, 81), (128, 41, 41), (16, 100, 100)]:
    C = 21
    image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, H, W), device=dev), 1)
    n0 = wseg_b200._lib.launch_count()
    ms = timed(wseg_b200.PAMR(10, D6).to(dev), image, mask)
    nl = (wseg_b200._lib.launch_count() - n0) // 23
    ms1 = timed(wseg_b200.PAMR(1, D6).to(dev), image, mask)
    print("B=%d %dx%d: PAMR forward %.4f ms (%d launches per call), %.1f Mpix/s; 1 iteration %.4f ms -> %.4f ms per further iteration"
          % (B, H, W, ms, nl, B * H * W / ms / 1e3, ms1, (ms - ms1) / 9), flush=True)
