"""Numerics study for the affinity kernel (run on CPU): fp32 emulation of candidate arithmetic against the oracle
(double-Welford std, IEEE divisions, accurate exp -- bit-equal to the reference on the goldens).

Candidates:
  two-pass : the round-1 kernel (centre-shifted two-pass std, Markstein divisions, accurate expf)
  one-pass : u = v - c once; S1 = sum u, S2 = sum u^2 (6 per-dilation partial sums); var = (S2 - S1^2/54)/53;
             a = sum_k |u| * (-(log2 e / K) * rcp(den_k))  (one FMA per tap and channel);
             w = ex2(a - max a) * rcp(sum)
Prints max-abs error of the 48 weights against the oracle for several input families."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import synth
from oracle import oracle

D6 = [1, 2, 4, 8, 12, 24]
f32 = np.float32


def fma(a, b, c):  # fp32 fused multiply-add (product exact in double; double rounding is negligible here)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def shifted(img, dy, dx):
    B, K, H, W = img.shape
    ys = np.clip(np.arange(H) + dy, 0, H - 1); xs = np.clip(np.arange(W) + dx, 0, W - 1)
    return img[:, :, ys][:, :, :, xs]


def one_pass(img, ulp_noise=None):
    B, K, H, W = img.shape
    taps = [(j // 3 - 1, j % 3 - 1) for j in range(9) if j != 4]
    u = np.empty((6, 8, B, K, H, W), f32)
    for i, d in enumerate(D6):
        for j, (ty, tx) in enumerate(taps):
            u[i, j] = shifted(img, ty * d, tx * d) - img
    s1 = np.zeros((6, B, K, H, W), f32); s2 = np.zeros((6, B, K, H, W), f32)
    for i in range(6):
        for j in range(8):
            s1[i] = s1[i] + u[i, j]
            s2[i] = fma(u[i, j], u[i, j], s2[i])
    S1 = ((s1[0] + s1[1]) + (s1[2] + s1[3])) + (s1[4] + s1[5])
    S2 = ((s2[0] + s2[1]) + (s2[2] + s2[3])) + (s2[4] + s2[5])
    m2 = fma(-(S1 * f32(1.0 / 54.0)), S1, S2)
    var = np.maximum(m2, f32(0)) * f32(1.0 / 53.0)
    sd = np.sqrt(var).astype(f32)
    den = f32(1e-8) + f32(0.1) * sd
    r = (f32(1) / den).astype(f32)
    nr = -(r * f32(1.4426950408889634 / K))
    a = np.zeros((6, 8, B, H, W), f32)
    for k in range(K):
        a = fma(np.abs(u[:, :, :, k]), np.broadcast_to(nr[:, k], a.shape).astype(f32), a)
    a = a.reshape(48, B, H, W)
    mx = a.max(0)
    e = np.exp2((a - mx).astype(np.float64)).astype(f32)
    if ulp_noise is not None:  # MUFU.EX2: 2 ulp
        e = (e * (1 + ulp_noise * 2.0 ** -23 * np.random.RandomState(0).uniform(-1, 1, e.shape))).astype(f32)
    s = np.zeros_like(mx)
    for p in range(48):
        s = s + e[p]
    w = e * (f32(1) / s).astype(f32)
    return np.moveaxis(w, 0, 1), sd


def families(B, H, W):
    rng = np.random.RandomState(7)
    yield "uniform", synth.image_uniform(B, 3, H, W, 0)
    yield "structured", synth.image_structured(B, 3, H, W, 1)
    yield "quantised", synth.image_structured(B, 3, H, W, 2, quantise=True)
    yield "constant", synth.image_constant(B, 3, H, W)
    flat = np.full((B, 3, H, W), 0.3, f32) + (1e-3 * rng.randn(B, 3, H, W)).astype(f32)
    yield "flat + 1e-3 noise", flat
    spikes = flat.copy(); m = rng.rand(B, 3, H, W) < 0.02; spikes[m] = 1.0
    yield "flat + bright outliers", spikes
    yield "bright 0.9..1 noise", (0.9 + 0.1 * rng.rand(B, 3, H, W)).astype(f32)
    yield "denormalised range 0..255", (255 * synth.image_structured(B, 3, H, W, 3)).astype(f32)
    ramp = np.broadcast_to(np.linspace(0, 1, W, dtype=f32), (B, 3, H, W)).copy()
    yield "ramp", ramp


if __name__ == "__main__":
    for name, img in families(2, 96, 120):
        ref = oracle.affinity(img, D6)
        sd_ref = oracle.local_std(img, D6)
        w, sd = one_pass(img, ulp_noise=2.0)
        rel = np.abs(sd - sd_ref) / np.maximum(sd_ref, 1e-30)
        print("%-28s one-pass: weights max-abs %.3g   std max-abs %.3g max-rel %.3g" %
              (name, np.abs(w - ref).max(), np.abs(sd - sd_ref).max(), rel[sd_ref > 1e-6].max() if (sd_ref > 1e-6).any() else 0))


def variant(img, std_mode="one", div_mode="fma", exp_noise=2.0, norm="rcp"):
    """Ablation: which step costs how much accuracy."""
    B, K, H, W = img.shape
    taps = [(j // 3 - 1, j % 3 - 1) for j in range(9) if j != 4]
    u = np.empty((6, 8, B, K, H, W), f32)
    for i, d in enumerate(D6):
        for j, (ty, tx) in enumerate(taps):
            u[i, j] = shifted(img, ty * d, tx * d) - img
    if std_mode == "exact":
        sd = oracle.local_std(img, D6)
    else:
        s1 = np.zeros((6, B, K, H, W), f32); s2 = np.zeros((6, B, K, H, W), f32)
        for i in range(6):
            for j in range(8):
                s1[i] = s1[i] + u[i, j]
        S1 = ((s1[0] + s1[1]) + (s1[2] + s1[3])) + (s1[4] + s1[5])
        mean = S1 * f32(1.0 / 54.0)
        if std_mode == "one":
            for i in range(6):
                for j in range(8):
                    s2[i] = fma(u[i, j], u[i, j], s2[i])
            S2 = ((s2[0] + s2[1]) + (s2[2] + s2[3])) + (s2[4] + s2[5])
            m2 = fma(-mean, S1, S2)
        else:  # two-pass
            for i in range(6):
                for j in range(8):
                    dv = u[i, j] - mean
                    s2[i] = fma(dv, dv, s2[i])
                s2[i] = fma(mean, mean, s2[i])  # the centre sample of this dilation (u = 0)
            m2 = ((s2[0] + s2[1]) + (s2[2] + s2[3])) + (s2[4] + s2[5])
        sd = np.sqrt(np.maximum(m2, f32(0)) * f32(1.0 / 53.0)).astype(f32)
    den = f32(1e-8) + f32(0.1) * sd
    if div_mode == "fma":
        nr = -((f32(1) / den).astype(f32) * f32(1.4426950408889634 / K))
        a = np.zeros((6, 8, B, H, W), f32)
        for k in range(K):
            a = fma(np.abs(u[:, :, :, k]), np.broadcast_to(nr[:, k], a.shape).astype(f32), a)
        a = a.reshape(48, B, H, W)
        mx = a.max(0)
        e = np.exp2((a - mx).astype(np.float64)).astype(f32)
    else:  # IEEE: per-channel division, sum, /K, exp
        a = np.zeros((6, 8, B, H, W), f32)
        for k in range(K):
            a = a + (-np.abs(u[:, :, :, k]) / den[:, k]).astype(f32)
        a = (a / f32(K)).astype(f32).reshape(48, B, H, W)
        mx = a.max(0)
        e = np.exp((a - mx).astype(np.float64)).astype(f32)
    e = (e * (1 + exp_noise * 2.0 ** -23 * np.random.RandomState(0).uniform(-1, 1, e.shape))).astype(f32)
    s = np.zeros_like(mx)
    for p in range(48):
        s = s + e[p]
    w = e * (f32(1) / s).astype(f32) if norm == "rcp" else (e / s).astype(f32)
    return np.moveaxis(w, 0, 1)


if __name__ == "__main__":
    print()
    for name, img in families(2, 96, 120):
        ref = oracle.affinity(img, D6)
        row = []
        for (sm, dm, en, nm) in [("one", "fma", 2.0, "rcp"), ("exact", "fma", 2.0, "rcp"), ("two", "fma", 2.0, "rcp"),
                                 ("two", "ieee", 0.5, "div"), ("two", "fma", 0.5, "rcp"), ("one", "ieee", 0.5, "div")]:
            row.append("%s/%s/%g/%s %.2e" % (sm, dm, en, nm, np.abs(variant(img, sm, dm, en, nm) - ref).max()))
        print("%-26s" % name, " | ".join(row))
