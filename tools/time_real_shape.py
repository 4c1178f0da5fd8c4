"""Time the pieces of the step at stage_net's real call shape (image 321x321, masks 81x81 or 41x41, B=16)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
from wseg_b200 import stage
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
B, C, H, W = 16, 21, 321, 321
pamr = wseg_b200.PAMR(10, D6).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(name, fn, n=50, cold=False):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    tot = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if cold:
        for _ in range(n):
            flush.zero_(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    else:
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); tot = e0.elapsed_time(e1)
    print("%-46s %.4f ms" % (name + (" [L2 flushed]" if cold else ""), tot / n), flush=True)
for h in (81, 41):
    print("masks %dx%d" % (h, h))
    image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, h, h), device=dev), 1)
    labels = (torch.rand((B, C - 1), device=dev) < 0.3).float(); labels[:, 0] = 1
    for cold in (False, True):
        t("refine_and_label", lambda: wseg_b200.refine_and_label(pamr, image, mask, labels), cold=cold)
        t("run_pamr (image resize + resident kernel)", lambda: stage.run_pamr(pamr, image, mask), cold=cold)
        ims = wseg_b200.resize_bilinear(image, (h, h)) if hasattr(wseg_b200, "resize_bilinear") else torch.nn.functional.interpolate(image, (h, h), mode="bilinear", align_corners=True)
        t("PAMR.forward on the resized image", lambda: pamr(ims, mask), cold=cold)
        dec = stage.run_pamr(pamr, image, mask)
        t("pseudo_labels (clean/max + labels, resize)", lambda: wseg_b200.pseudo_labels(dec, labels, (H, W)), cold=cold)
        t("rescale_and_clean with class max", lambda: stage.rescale_and_clean(dec, (H, W), labels, return_class_max=True), cold=cold)
# host-side enqueue time of one call (no synchronisation inside the loop): when it exceeds the device time, the step is launch-bound
import time
for h in (81, 41):
    image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, h, h), device=dev), 1)
    labels = (torch.rand((B, C - 1), device=dev) < 0.3).float(); labels[:, 0] = 1
    for _ in range(10): wseg_b200.refine_and_label(pamr, image, mask, labels)
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n): wseg_b200.refine_and_label(pamr, image, mask, labels)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("masks %dx%d: host enqueue %.4f ms per refine_and_label, with the final synchronize %.4f ms" % (h, h, (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3))
