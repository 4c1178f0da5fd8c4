"""Repeat one forward many times and count outputs that differ bitwise from the first (race detector)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
H = int(sys.argv[1]) if len(sys.argv) > 1 else 321
W = int(sys.argv[2]) if len(sys.argv) > 2 else 321
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
N = int(sys.argv[4]) if len(sys.argv) > 4 else 200
image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, 21, H, W), device=dev), 1)
pamr = wseg_b200.PAMR(int(os.environ.get("DET_ITERS", "10")), D6).to(dev)
ref = pamr(image, mask).clone()
bad = 0; worst = 0.0; where = set()
for i in range(N):
    junk = torch.rand((B, 21, H, W), device=dev)   # allocator / L2 churn between calls, like the tests
    out = pamr(image, 0.5 * mask + 0.5 * mask)
    if not torch.equal(out, ref):
        bad += 1
        d = (out - ref).abs()
        worst = max(worst, float(d.max()))
        idx = torch.nonzero(d.flatten(2).max(-1).values.max(0).values > 0).flatten().tolist()
        ys = torch.nonzero(d.amax((0, 1)).amax(1) > 0).flatten().tolist()
        xs = torch.nonzero(d.amax((0, 1)).amax(0) > 0).flatten().tolist()
        where.add((min(ys), max(ys), min(xs), max(xs)))
        if bad <= 3:
            nz = torch.nonzero(d > 0)
            tiles = {}
            for b, c, y, x in nz[:: max(1, len(nz) // 4000)].tolist():
                tiles.setdefault((b, y // 40, x // 32), set()).add(c)
            print("run %d: %d differing values; (b,ty,tx) -> classes:" % (i, len(nz)),
                  {k: sorted(v) for k, v in sorted(tiles.items())[:12]}, flush=True)
            if len(nz) < 400:
                print("   pixels (b,c,y,x):", nz.tolist()[:40], "values", out[tuple(nz[0].tolist())].item(), ref[tuple(nz[0].tolist())].item(), flush=True)
print("H=%d W=%d B=%d: %d of %d runs differ, worst %.3g, bounding boxes (y0,y1,x0,x1): %s  env %s" % (
    H, W, B, bad, N, worst, sorted(where)[:6], {k: v for k, v in os.environ.items() if k.startswith("PAMR_B200")}))
