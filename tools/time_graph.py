"""Eager vs CUDA-graph replay of refine_and_label at stage_net's real call shapes (mask below image resolution)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
dev = "cuda:0"
D6 = [1, 2, 4, 8, 12, 24]
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
pamr = wseg_b200.PAMR(10, D6).to(dev)
for (B, h, w, H, W) in [(16, 41, 41, 321, 321), (16, 81, 81, 321, 321), (16, 321, 321, 321, 321), (1, 321, 321, 321, 321)]:
    img = torch.rand((B, 3, H, W), device=dev); msk = torch.softmax(2 * torch.randn((B, 21, h, w), device=dev), 1)
    lab = (torch.rand((B, 20), device=dev) < 0.3).float()
    eager = lambda: wseg_b200.refine_and_label(pamr, img, msk, lab)
    ms_e = t(eager)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = wseg_b200.refine_and_label(pamr, img, msk, lab)
    ms_g = t(g.replay)
    print("B=%d mask %dx%d image %dx%d: eager %.3f ms, graph replay %.3f ms (%.2fx)" % (B, h, w, H, W, ms_e, ms_g, ms_e / ms_g))
