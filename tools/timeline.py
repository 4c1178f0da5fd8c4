"""Print the per-pass timeline of CTA 0 of one tuned propagation launch (debug hook)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
B, C, H, W = 16, 21, int(os.environ.get("PROF_H", 320)), int(os.environ.get("PROF_W", 320))
image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, H, W), device=dev), 1)
aff = wseg_b200.local_affinity(image, D6)
for _ in range(3): wseg_b200.propagate(aff, mask, D6, 1)
buf = torch.zeros((2, 4096, 2), dtype=torch.int64, device=dev)
raw = ctypes.CDLL(wseg_b200._lib.LIB_PATH)
raw.pamr_debug_set_timeline.argtypes = [ctypes.c_void_p]
torch.cuda.synchronize()
raw.pamr_debug_set_timeline(buf.data_ptr())
wseg_b200.propagate(aff, mask, D6, 1)
torch.cuda.synchronize()
raw.pamr_debug_set_timeline(None)
ev = buf.cpu().numpy()
t0 = min(ev[g, 0, 0] for g in range(2))
names = {1: "tile-begin", 2: "bar1-done", 3: "fill-done", 4: "bar2-done", 6: "wait-done", 7: "compute-done", 8: "store-done"}
for g in range(2):
    print("---- group", g)
    prev = None; ntile = 0
    for t, code in ev[g]:
        if t == 0: break
        if code == 1: ntile += 1
        if ntile in (3, 4):
            nm = names.get(int(code), "pass %d begin" % (code - 100))
            print("%10d  +%7d  %s" % (t - t0, 0 if prev is None else t - prev, nm))
        prev = t
