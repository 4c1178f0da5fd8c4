"""Does a small kernel on a second stream run WHILE the persistent tile kernel is resident?
Stream A: one tuned propagation launch (320x320, B=16 -> ~250 us).  Stream B: N small torch
kernels.  Prints when B's work finished relative to A's start / end."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
B, C, H, W = 16, 21, 320, 320
image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, H, W), device=dev), 1)
aff = wseg_b200.local_affinity(image, D6)
x = torch.rand((1 << 16,), device=dev); y = torch.empty_like(x)
sB = torch.cuda.Stream()
for trial in range(3):
    out = wseg_b200.propagate(aff, mask, D6, 1)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    e[0].record()
    sB.wait_event(e[0])
    out = wseg_b200.propagate(aff, mask, D6, 4)   # relayout + 4 tile launches on the current stream
    e[1].record()
    with torch.cuda.stream(sB):
        e[2].record(sB)
        for _ in range(20):
            torch.add(x, 1.0, out=y)
        e[3].record(sB)
    torch.cuda.synchronize()
    print("A: %.1f us total;  B: started at %.1f us, finished at %.1f us" % (
        1e3 * e[0].elapsed_time(e[1]), 1e3 * e[0].elapsed_time(e[2]), 1e3 * e[0].elapsed_time(e[3])))
