"""Seeded synthetic input families for the PAMR hot path.

numpy-only so that the same inputs are produced here, on the GPU box and in
oracle/gen_golden.py, independent of the torch version.
"""
import numpy as np


def image_uniform(B, K, H, W, seed=0):
    """(i) uniform [0,1) noise."""
    return np.random.default_rng(seed).random((B, K, H, W), dtype=np.float32)


def image_structured(B, K, H, W, seed=1, quantise=False):
    """(ii) constant base + 12 random axis-aligned rectangles + N(0,0.02) noise, clamped to [0,1];
    (iii) the same rounded to k/255 (exercises exact zero-std / zero-difference paths)."""
    rng = np.random.default_rng(seed)
    img = np.empty((B, K, H, W), dtype=np.float32)
    for b in range(B):
        img[b] = rng.random((K, 1, 1), dtype=np.float32)
        for _ in range(12):
            y0, y1 = sorted(rng.integers(0, H + 1, 2))
            x0, x1 = sorted(rng.integers(0, W + 1, 2))
            img[b, :, y0:y1, x0:x1] = rng.random((K, 1, 1), dtype=np.float32)
    if quantise:
        # flat regions stay exactly flat: no noise, round to 8 bit
        return (np.round(np.clip(img, 0, 1) * 255.0) / 255.0).astype(np.float32)
    img += rng.normal(0.0, 0.02, img.shape).astype(np.float32)
    return np.clip(img, 0.0, 1.0).astype(np.float32)


def image_constant(B, K, H, W, value=0.5):
    """(iv) constant image: sigma = 0 -> uniform 1/P weights."""
    return np.full((B, K, H, W), value, dtype=np.float32)


def _softmax(x, axis):
    x = x - x.max(axis=axis, keepdims=True)
    e = np.exp(x)
    return (e / e.sum(axis=axis, keepdims=True)).astype(np.float32)


def mask_softmax(B, C, H, W, seed=10, temp=2.0):
    """softmax(temp * randn) over classes."""
    rng = np.random.default_rng(seed)
    return _softmax(temp * rng.standard_normal((B, C, H, W)).astype(np.float32), 1)


def mask_blobs(B, C, H, W, seed=11):
    """CAM-like smooth masks: softmax of a few Gaussian bumps per class (confident regions,
    so that the pseudo-label thresholds select non-trivial areas)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    logits = np.zeros((B, C, H, W), dtype=np.float32)
    for b in range(B):
        for c in range(C):
            for _ in range(2):
                cy, cx = rng.random() * H, rng.random() * W
                s = (0.08 + 0.25 * rng.random()) * max(H, W)
                logits[b, c] += 4.0 * rng.random() * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
    logits[:, 0] += 1.0
    return _softmax(logits, 1)


def labels_bernoulli(B, C, seed=20, p=0.15):
    """[B, C-1] float 0/1 image-level labels, column 0 forced to 1."""
    rng = np.random.default_rng(seed)
    lab = (rng.random((B, C - 1)) < p).astype(np.float32)
    lab[:, 0] = 1.0
    return lab
