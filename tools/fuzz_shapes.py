"""Random-shape fuzz of the forward path against the CPU oracle (tail / lane / strip / generic paths):
python tools/fuzz_shapes.py [N] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, wseg_b200
from oracle import oracle
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
D6 = [1, 2, 4, 8, 12, 24]
worst = 0.0
t0 = time.time()
for n in range(N):
    B, C = int(rng.randint(1, 7)), int(rng.randint(1, 25))
    H, W = int(rng.randint(8, 200)), int(rng.randint(8, 200))
    if rng.rand() < 0.3:  # remainders of one tile size or other
        H, W = 40 * int(rng.randint(1, 5)) + int(rng.randint(0, 9)), 32 * int(rng.randint(1, 6)) + int(rng.randint(0, 9))
    it = int(rng.choice([1, 2, 5]))
    image = rng.rand(B, 3, H, W).astype(np.float32)
    e = np.exp(rng.randn(B, C, H, W)).astype(np.float32); mask = e / e.sum(1, keepdims=True)
    out = wseg_b200.PAMR(it, D6).cuda()(torch.from_numpy(image).cuda(), torch.from_numpy(mask).cuda()).cpu().numpy()
    err = float(np.abs(out - oracle.pamr_forward(image, mask, it, D6)).max())
    worst = max(worst, err)
    if err > 1e-5:
        print("MISMATCH B=%d C=%d H=%d W=%d iters=%d: %.3g" % (B, C, H, W, it, err)); sys.exit(1)
print("fuzz ok: %d shapes, worst max-abs %.3g, %.1f s" % (N, worst, time.time() - t0))
