"""Print the per-pass timeline of one CTA of one tuned propagation launch (experiment builds only:
tools/build_variant.sh NAME pamr_propagate_sm100.cu -DPAMR_EXPERIMENTS, then PAMR_LIB=... python tools/timeline.py)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
B, C, H, W = 16, 21, int(os.environ.get("PROF_H", 320)), int(os.environ.get("PROF_W", 320))
CTA = int(os.environ.get("PROF_CTA", 0))
image = torch.rand((B, 3, H, W), device=dev); mask = torch.softmax(2 * torch.randn((B, C, H, W), device=dev), 1)
aff = wseg_b200.local_affinity(image, D6)
for _ in range(3): wseg_b200.propagate(aff, mask, D6, 2)
buf = torch.zeros((5, 4096, 2), dtype=torch.int64, device=dev)
raw = ctypes.CDLL(wseg_b200._lib.LIB_PATH)
raw.pamr_debug_set_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
torch.cuda.synchronize()
raw.pamr_debug_set_timeline(buf.data_ptr(), CTA, int(os.environ.get("PROF_SKIP", 1)))   # record the 2nd of 3 iterations (row-pair in, row-pair out); fused launch: PROF_SKIP=0
wseg_b200.propagate(aff, mask, D6, int(os.environ.get("PROF_ITERS", 3)))
torch.cuda.synchronize()
ev = buf.cpu().numpy()
t0 = min(int(ev[g, 0, 0]) for g in range(3) if ev[g, 0, 0] > 0)
issue = {}
names = {6: "wait-done", 9: "acc-ready", 8: "store-done"}
TILES = [int(x) for x in os.environ.get("PROF_TILES", "3,4").split(",")]
for g in range(3):
    print("---- group", g)
    prev = None; ntile = 0; lastk = 99
    for t, code in ev[g]:
        if t == 0: break
        t, code = int(t), int(code)
        if code >= 100:
            k = code - 100
            if k < lastk: ntile += 1
            lastk = k
        if ntile in TILES:
            nm = names.get(code, "pass %d begin" % (code - 100))
            if code >= 100:
                nn = (ntile - 1) * C + (code - 100)
                if nn in issue: nm += "   (TMA issued %d cycles earlier)" % (t - issue[nn])
            print("%10d  +%7d  %s" % (t - t0, 0 if prev is None else t - prev, nm))
        prev = t
for st, nm in ((3, "loader (2000+u: bulk copy of unit u issued)"), (4, "issuer (2100+u: unit u staged and free, tcgen05.cp issued)")):
    print("----", nm)
    prev = None; ntile = 0; lastu = 99
    for t, code in ev[st]:
        if t == 0: break
        t, code = int(t), int(code)
        u = code % 100
        if u < lastu: ntile += 1
        lastu = u
        if ntile in TILES or ntile - 1 in TILES:
            print("%10d  +%7d  tile %d  %d" % (t - t0, 0 if prev is None else t - prev, ntile, code))
        prev = t
# per-tile summary per group
for g in range(3):
    rows = [(int(t), int(c)) for t, c in ev[g] if t != 0]
    starts = []; lastk = 99
    for t, c in rows:
        if c >= 100:
            if c - 100 < lastk: starts.append(t)
            lastk = c - 100
    print("group %d tile durations:" % g, [b - a for a, b in zip(starts, starts[1:])][:12])
