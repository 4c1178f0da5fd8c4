"""Time the bench step and its pieces at config 2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
B, C, H, W = 16, 21, 321, 321
image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, H, W), device=dev), 1)
labels = (torch.rand((B, C - 1), device=dev) < 0.3).float(); labels[:, 0] = 1
pamr = wseg_b200.PAMR(10, D6).to(dev)
def t(name, fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-40s %.3f ms" % (name, e0.elapsed_time(e1) / n), flush=True)
t("forward", lambda: pamr(image, mask))
t("forward + class max", lambda: pamr(image, mask, return_class_max=True))
t("refine_and_label", lambda: wseg_b200.refine_and_label(pamr, image, mask, labels))
dec, cmax = pamr(image, mask, return_class_max=True)
t("pseudo_labels (fused max)", lambda: wseg_b200.pseudo_labels(dec, labels, None, cmax))
t("pseudo_labels (own max)", lambda: wseg_b200.pseudo_labels(dec, labels, None, None))
