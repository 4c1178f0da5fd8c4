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
buf = torch.zeros((3, 4096, 2), dtype=torch.int64, device=dev)
raw = ctypes.CDLL(wseg_b200._lib.LIB_PATH)
raw.pamr_debug_set_timeline.argtypes = [ctypes.c_void_p]
torch.cuda.synchronize()
raw.pamr_debug_set_timeline(buf.data_ptr())
wseg_b200.propagate(aff, mask, D6, 1)
torch.cuda.synchronize()
raw.pamr_debug_set_timeline(None)
ev = buf.cpu().numpy()
t0 = min(ev[g, 0, 0] for g in range(2))
issue = {int(c) - 1000: int(t) for t, c in ev[2] if c >= 1000}
names = {5: "tma-landed", 1: "tile-begin", 2: "bar1-done", 3: "fill-done", 4: "bar2-done", 6: "wait-done", 7: "syncwarp-done", 9: "acc-ready", 8: "store-done"}
for g in range(2):
    print("---- group", g)
    prev = None; ntile = 0
    for t, code in ev[g]:
        if t == 0: break
        if code == 1: ntile += 1
        if ntile in (3, 4):
            nm = names.get(int(code), "pass %d begin" % (code - 100))
            if code >= 100:
                n = (ntile - 1) * C + int(code - 100)
                nm += "   (TMA for this class issued at %d = %d cycles earlier)" % (issue.get(n, 0) - t0, t - issue.get(n, 0))
            print("%10d  +%7d  %s" % (t - t0, 0 if prev is None else t - prev, nm))
        prev = t

# per-tile summary for group 0: fill, blocking waits, compute
import collections
g = 0
rows = [(int(t), int(c)) for t, c in ev[g] if t != 0]
tiles = []
cur = None
for i, (t, c) in enumerate(rows):
    if c == 1:
        cur = {"begin": t, "wait": 0, "fill": 0, "n": 0}
        tiles.append(cur)
    elif c == 2: cur["b1"] = t
    elif c == 3: cur["fill"] = t - cur["b1"]
    elif c >= 100: cur["pb"] = t
    elif c == 6: cur["wait"] += t - cur["pb"]; cur["n"] += 1
    elif c == 8: cur["end"] = t
print("tile summaries (group 0): total / fill / sum of waits")
for tl in tiles[:10]:
    if "end" in tl: print("  total %7d  fill %6d  waits %6d over %d passes" % (tl["end"] - tl["begin"], tl["fill"], tl["wait"], tl["n"]))
