import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
def t(name, fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-50s %.3f ms" % (name, e0.elapsed_time(e1) / n), flush=True)
for (H, W) in [(324, 320), (320, 321)]:
    B, C = 16, 21
    image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, H, W), device=dev), 1)
    aff = wseg_b200.local_affinity(image, D6)
    for it in (1, 2):
        t("%dx%d propagate iters=%d" % (H, W, it), lambda: wseg_b200.propagate(aff, mask, D6, it))
        t("%dx%d propagate iters=%d + class max" % (H, W, it), lambda: wseg_b200.propagate(aff, mask, D6, it, return_class_max=True))
