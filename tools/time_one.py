"""Time the forward path pieces at one shape: python tools/time_one.py [H W [B]]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
H = int(sys.argv[1]) if len(sys.argv) > 1 else 320
W = int(sys.argv[2]) if len(sys.argv) > 2 else 320
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
C = 21
image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, H, W), device=dev), 1)
pamr = wseg_b200.PAMR(10, D6).to(dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
pamr2 = wseg_b200.PAMR(20, D6).to(dev)
fw = t(lambda: pamr(image, mask))
fw2 = t(lambda: pamr2(image, mask))
print("H=%d W=%d B=%d  marginal launch %.4f ms = %.1f%% of the HBM roofline (360 B/px at 6543.1 GB/s); forward(10) %.3f ms" % (
    H, W, B, (fw2 - fw) / 10, 100 * 360e-9 * B * H * W / ((fw2 - fw) / 10 * 1e-3) / 6543.1, fw), flush=True)
af = t(lambda: wseg_b200.local_affinity(image, D6))
print("H=%d W=%d B=%d  forward %.3f ms (affinity-std alone %.3f ms) -> (fwd-aff)/10 = %.3f ms/iter ; tags %s %s" % (
    H, W, B, fw, af, (fw - af) / 10, os.environ.get("PAMR_B200_STAGGER_CTA", "-"), os.environ.get("PAMR_B200_STAGGER_GRP", "-")), flush=True)
