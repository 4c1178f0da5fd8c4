"""Times PAMR forward (+ epilogue) on the BASELINE.json shape families (configs 2-5, one GPU) and
checks size-independent properties on the result.  Output is committed under profiles/."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, wseg_b200
D6 = [1, 2, 4, 8, 12, 24]
dev = "cuda:0"
PEAK = 6543.1
def t(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
pamr = wseg_b200.PAMR(10, D6).to(dev)
cases = [("config 2: B=16 321x321", 16, 321, 321), ("config 3: B=1 256x256", 1, 256, 256), ("config 3: B=1 512x512", 1, 512, 512),
         ("config 3: B=1 768x768", 1, 768, 768), ("config 3: B=1 1024x1024", 1, 1024, 1024),
         ("config 4: B=8 1024x2048", 8, 1024, 2048), ("config 5 (1 GPU): B=128 321x321", 128, 321, 321),
         ("stage_net real shape: B=16 81x81", 16, 81, 81), ("stage_net real shape: B=16 41x41", 16, 41, 41)]
for name, B, H, W in cases:
    C = 21
    g = torch.Generator(device=dev).manual_seed(1)
    image = torch.rand((B, 3, H, W), generator=g, device=dev)
    mask = torch.softmax(2 * torch.randn((B, C, H, W), generator=g, device=dev), 1)
    labels = (torch.rand((B, C - 1), generator=g, device=dev) < 0.3).float(); labels[:, 0] = 1
    ms_f = t(lambda: pamr(image, mask))
    ms_s = t(lambda: wseg_b200.refine_and_label(pamr, image, mask, labels))
    out = pamr(image, mask)
    lab = wseg_b200.refine_and_label(pamr, image, mask, labels)
    ok_sum = float((out.sum(1) - 1).abs().max())
    ok_rng = bool((out.flatten(2).max(-1).values <= mask.flatten(2).max(-1).values + 1e-6).all())
    npx = B * H * W
    print("%-36s PAMR %8.3f ms = %7.1f Mpix/s = %6.1f GB/s algorithmic (%4.1f%% of %.0f) | +epilogue %8.3f ms = %7.1f Mpix/s | "
          "|sum_c-1| %.1e convex %s labels in {0..20,255}: %s" % (
              name, ms_f, npx / ms_f / 1e3, 3804 * npx / ms_f / 1e6, 3804 * npx / ms_f / 1e6 / PEAK * 100, PEAK, ms_s,
              npx / ms_s / 1e3, ok_sum, ok_rng, set(torch.unique(lab).tolist()) <= set(range(21)) | {255}), flush=True)
    del image, mask, out, lab
    torch.cuda.empty_cache()
